#!/usr/bin/env bash
# Round 2, GPU call 33 (1 GPU): the final tree - whole GPU suite, smoke, the default bench line as the driver runs it
set -u
mkdir -p gpurun_out
S=$SECONDS
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r2ag_tests.log
tail -3 gpurun_out/r2ag_tests.log
echo "== smoke ==" >> gpurun_out/r2ag_tests.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2 >> gpurun_out/r2ag_tests.log
tail -1 gpurun_out/r2ag_tests.log
echo "tests: $((SECONDS-S)) s"; S=$SECONDS
timeout 600 python bench.py > gpurun_out/r2ag_bench_default.json 2> gpurun_out/r2ag_bench_default.err
echo "own arm: $((SECONDS-S)) s"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2ag_bench_default.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], 'parity', d['parity']['ok'])
print('clocks', d.get('clocks'), 'launches', d.get('gpu_launches'))
mt = d.get('model_train') or {}
print('model_train', mt.get('value'), mt.get('ms_per_step'), mt.get('execution'))
mp = d.get('model_train_partitioned') or {}
print('partitioned', json.dumps({k: mp.get(k) for k in ('value', 'ms_per_step', 'execution')})[:300])
for c, v in (d.get('named_configs') or {}).items():
    print(c, v.get('value'), v.get('ms_per_step'), (v.get('parity') or {}).get('ok'))
PY
