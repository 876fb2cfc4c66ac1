#!/usr/bin/env bash
# Round 2, GPU call 28 (1 GPU): deeper cp.async pipeline of the weight-gradient kernel for narrow operands, patch kernel
# with 16 warps - the tests of those kernels, the HealpyGCNN step, the layer bench (headline kernel must not move)
set -u
mkdir -p gpurun_out
S=$SECONDS
timeout 900 python -m pytest tests/test_gpu_tensor_core.py tests/test_gpu_lattice.py tests/test_gpu_model.py tests/test_gpu_r2_networks.py -q -m gpu -x 2>&1 | tail -8 > gpurun_out/r2ab_tests.log
tail -3 gpurun_out/r2ab_tests.log
echo "tests: $((SECONDS-S)) s"; S=$SECONDS
timeout 600 python bench.py --model-only > gpurun_out/r2ab_model.json 2> gpurun_out/r2ab_model.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2ab_model.json').read().strip().splitlines()[-1])
print('model_train', d.get('value'), d.get('ms_per_step'), d.get('eager_ms_per_step'), d.get('execution'), (d.get('cuda_graph') or {}).get('validated'))
PY
timeout 600 python bench.py --no-e2e --no-cpu-baseline --no-other-modes --no-f-sweep --no-partitioned --no-configs --no-model > gpurun_out/r2ab_layer.json 2> gpurun_out/r2ab_layer.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2ab_layer.json').read().strip().splitlines()[-1])
print('layer', d['value'], d['ms_per_step'], d['parity']['ok'], d['roofline']['frac'])
PY
echo "bench: $((SECONDS-S)) s"; S=$SECONDS
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2ab_launches_model_train.csv \
  python bench.py --model-only --no-graph --steps 1 --warmup 1 > gpurun_out/r2ab_ncu_model.log 2>&1
echo "ncu: $((SECONDS-S)) s"
