#!/usr/bin/env bash
# Round 2, GPU call 35 (1 GPU, the last 2 minutes): weight-gradient kernel with two 16-row sub-stages per pipeline stage for
# narrow operands - its tests, then the HealpyGCNN step with and without (DEEPSPHERE_TN_HALVES=1)
set -u
mkdir -p gpurun_out
timeout 80 python -m pytest tests/test_gpu_tensor_core.py -q -m gpu -x 2>&1 | tail -2 > gpurun_out/r2ai_tests.log
tail -1 gpurun_out/r2ai_tests.log
for h in 2 1; do
  DEEPSPHERE_TN_HALVES=$h timeout 60 python bench.py --model-only --no-graph > gpurun_out/r2ai_model_halves$h.json 2> gpurun_out/r2ai_model_halves$h.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r2ai_model_halves$h.json').read().strip().splitlines()[-1])
print('halves', $h, d.get('value'), d.get('ms_per_step'), d.get('final_loss'), d.get('first_step',{}).get('loss'))"
done
