#!/usr/bin/env bash
# Round 2, GPU call 23: row-per-thread SpMM for narrow F - GPU suite, named configs
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 > gpurun_out/r2w_tests.log
tail -4 gpurun_out/r2w_tests.log
timeout 1500 python bench.py --no-e2e --no-other-modes --no-f-sweep --no-partitioned > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2w_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'parity ok', d['parity']['ok'], 'model_train', d['model_train']['ms_per_step'])
for c, v in (d.get('named_configs') or {}).items():
    print(c, json.dumps({a: v.get(a) for a in ('value','ms_per_step','eager_ms_per_step','execution','error')})[:300], json.dumps((v.get('parity') or {}).get('rel_err')))
PY
