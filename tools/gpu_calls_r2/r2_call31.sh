#!/usr/bin/env bash
# Round 2, GPU call 31 (1 GPU): weight-gradient kernel with 6 / 10 stages in flight for narrow operands - its tests, the
# HealpyGCNN step
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensor_core.py tests/test_gpu_model.py -q -m gpu -x 2>&1 | tail -5 > gpurun_out/r2ae_tests.log
tail -2 gpurun_out/r2ae_tests.log
timeout 600 python bench.py --model-only > gpurun_out/r2ae_model.json 2> gpurun_out/r2ae_model.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2ae_model.json').read().strip().splitlines()[-1])
print('model_train', d.get('value'), d.get('ms_per_step'), d.get('eager_ms_per_step'), d.get('execution'), (d.get('cuda_graph') or {}).get('validated'))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2ae_launches_model_train.csv \
  python bench.py --model-only --no-graph --steps 1 --warmup 1 > gpurun_out/r2ae_ncu_model.log 2>&1
