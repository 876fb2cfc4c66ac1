#!/usr/bin/env bash
# Round 2, GPU call 27 (1 GPU): patch kernel for the irregular rows (ds_patch.cu), parallel split-K reduction, fused Adam -
# GPU suite, then the HealpyGCNN step A/B (patch kernel on / off, Adam fused / foreach) and its launch list
set -u
mkdir -p gpurun_out
S=$SECONDS
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2aa_tests.log
tail -3 gpurun_out/r2aa_tests.log
echo "tests: $((SECONDS-S)) s"; S=$SECONDS
for v in "1 fused" "0 fused" "1 foreach"; do
  set -- $v
  DEEPSPHERE_PATCH=$1 timeout 600 python bench.py --model-only --adam $2 > gpurun_out/r2aa_model_patch$1_$2.json 2> gpurun_out/r2aa_model_patch$1_$2.err
  python - "$1" "$2" <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/r2aa_model_patch{sys.argv[1]}_{sys.argv[2]}.json').read().strip().splitlines()[-1])
mt = d.get('model_train') or d
print('patch', sys.argv[1], 'adam', sys.argv[2], mt.get('value'), mt.get('ms_per_step'), mt.get('eager_ms_per_step'), mt.get('execution'), (mt.get('cuda_graph') or {}).get('validated'))
PY
done
echo "model A/B: $((SECONDS-S)) s"; S=$SECONDS
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2aa_launches_model_train.csv \
  python bench.py --model-only --no-graph --steps 3 --warmup 1 > gpurun_out/r2aa_ncu_model.log 2>&1
echo "ncu: $((SECONDS-S)) s"
