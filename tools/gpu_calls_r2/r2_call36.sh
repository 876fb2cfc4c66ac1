#!/usr/bin/env bash
# Round 2, GPU call 36 (1 GPU, the last minute): weight-gradient kernel with the proxy fence on the issuing warp instead of
# the producers (whose MEMBAR waits for every copy in flight) - its tests, then the HealpyGCNN step both ways
set -u
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_tensor_core.py -q -m gpu -x 2>&1 | tail -2 > gpurun_out/r2aj_tests.log
tail -1 gpurun_out/r2aj_tests.log
for f in issuer producer; do
  DEEPSPHERE_TN_FENCE=$f timeout 40 python bench.py --model-only --no-graph > gpurun_out/r2aj_model_fence_$f.json 2> gpurun_out/r2aj_model_fence_$f.err
  python -c "
import json
d=json.loads(open('gpurun_out/r2aj_model_fence_$f.json').read().strip().splitlines()[-1])
print('fence', '$f', d.get('value'), d.get('ms_per_step'), d.get('final_loss'), d.get('first_step',{}).get('loss'), d.get('first_step',{}).get('grad_norms')[:4])"
done
