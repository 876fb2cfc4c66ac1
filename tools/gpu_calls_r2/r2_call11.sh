#!/usr/bin/env bash
# Round 2, GPU call 11: IO-side accumulator drain variants; streaming pseudo-convolution kernels per layer
set -u
mkdir -p gpurun_out
for name in iodrain iodrainsymw iodrainio48; do
  V=$PWD/deepsphere-cosmo-tf2_b200/lib/libdeepsphere_b200_$name.so
  [ -f "$V" ] || continue
  {
    DEEPSPHERE_LIB=$V timeout 300 python -m pytest tests/test_gpu_lattice.py -q -m gpu -k "fused or conv2" 2>&1 | tail -3
    DEEPSPHERE_LIB=$V timeout 200 python tools/bench_fwd.py tf32 32 2>&1 | grep RESULT
    DEEPSPHERE_LIB=$V timeout 200 python tools/bench_fwd.py tf32 32 bwd 2>&1 | grep RESULT
  } > gpurun_out/r2k_variant_$name.log 2>&1
  echo "$name: $(grep -E 'passed|failed|rror' gpurun_out/r2k_variant_$name.log | tail -1) | $(grep RESULT gpurun_out/r2k_variant_$name.log | cut -c1-100 | tr '\n' '|')"
done
for sk in 0 1; do
  DEEPSPHERE_SKINNY=$sk timeout 600 python tools/part_layer_times.py 8 2>/dev/null | grep '^{' > gpurun_out/r2k_layers_skinny$sk.json
  DEEPSPHERE_SKINNY=$sk timeout 300 python bench.py --model-only --no-graph 2>/dev/null | grep '^{' > gpurun_out/r2k_model_skinny$sk.json
done
python - <<'PY'
import json
a = json.load(open('gpurun_out/r2k_layers_skinny0.json')); b = json.load(open('gpurun_out/r2k_layers_skinny1.json'))
print('skinny 0 -> 1: fwd', a['fwd_total'], b['fwd_total'], 'bwd', a['bwd_total'], b['bwd_total'])
for x, y in zip(a['layers'], b['layers']):
    print(f"{x['layer']:28s} fwd {x['fwd_ms']:8.3f} -> {y['fwd_ms']:8.3f}   bwd {x['bwd_ms']:8.3f} -> {y['bwd_ms']:8.3f}")
for sk in (0, 1):
    d = json.load(open(f'gpurun_out/r2k_model_skinny{sk}.json'))
    print('model_train skinny', sk, d['ms_per_step'], d['first_step']['loss'])
PY
