#!/usr/bin/env bash
# Round 2, GPU call 17 (2 GPUs): whole GPU suite incl. the 2-GPU partition tests (BatchNorm), smoke, default bench at N = 1
set -u
mkdir -p gpurun_out
{
  timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -12
  echo "== smoke =="
  CUDA_VISIBLE_DEVICES=0 timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
} > gpurun_out/r2q_tests.log 2>&1
tail -8 gpurun_out/r2q_tests.log
SECONDS=0; CUDA_VISIBLE_DEVICES=0 timeout 1500 python bench.py > gpurun_out/r2q_bench_default.json 2> gpurun_out/r2q_bench_default.err
echo "bench wall seconds: $SECONDS"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/r2q_bench_default.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('value', d['value'], 'ms', d['ms_per_step'], 'parity', json.dumps(d['parity']['rel_err']), 'launches', d['gpu_launches'])
    for k in ('model_train','model_train_partitioned'):
        m = d.get(k) or {}
        print(k, json.dumps({a: m.get(a) for a in ('value','ms_per_step','eager_ms_per_step','global_batch','peak_memory_GB','execution','error')})[:500])
    for c, v in (d.get('named_configs') or {}).items():
        print(c, json.dumps({a: v.get(a) for a in ('value','ms_per_step','eager_ms_per_step','execution','error')})[:300], json.dumps((v.get('parity') or {}).get('rel_err')))
except Exception as e:
    print('no json', e)
PY
