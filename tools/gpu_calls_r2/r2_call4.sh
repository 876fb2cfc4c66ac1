#!/usr/bin/env bash
# Round 2, GPU call 4: new tests (shapes, BatchNorm kernels, residual values, halo kernels), whole suite, default bench
set -u
mkdir -p gpurun_out
{
  timeout 900 python -m pytest tests/test_gpu_r2_shapes.py tests/test_gpu_partition.py -q -m gpu 2>&1 | tail -25
  echo "== whole GPU suite =="
  timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
} > gpurun_out/r2d_tests.log 2>&1
tail -12 gpurun_out/r2d_tests.log
timeout 1200 python bench.py > gpurun_out/r2d_bench_default.json 2> gpurun_out/r2d_bench_default.err
tail -c 600 gpurun_out/r2d_bench_default.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2d_bench_default.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','roofline','parity','e2e','model_train_partitioned','partition_parity'):
        print(k, json.dumps(d.get(k))[:900])
    print('model_train', json.dumps({k:v for k,v in (d.get('model_train') or {}).items() if k in ('value','ms_per_step','error')}))
except Exception as e:
    print('no json', e)
PY
