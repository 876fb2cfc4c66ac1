#!/usr/bin/env bash
# Round 2, GPU call 22: narrow contraction kernels (ds_narrow.cu) - whole GPU suite, then the named configs
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 > gpurun_out/r2v_tests.log
tail -6 gpurun_out/r2v_tests.log
SECONDS=0; timeout 1500 python bench.py --no-e2e --no-other-modes --no-f-sweep > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
echo "bench wall seconds: $SECONDS"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2v_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'parity ok', d['parity']['ok'])
print('model_train', d['model_train']['ms_per_step'], 'partitioned', d['model_train_partitioned']['ms_per_step'])
for c, v in (d.get('named_configs') or {}).items():
    print(c, json.dumps({a: v.get(a) for a in ('value','ms_per_step','eager_ms_per_step','execution','error')})[:300], json.dumps((v.get('parity') or {}).get('rel_err')), json.dumps((v.get('cpu_baseline') or {}).get('value')))
PY
