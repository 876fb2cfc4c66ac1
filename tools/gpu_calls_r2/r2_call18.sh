#!/usr/bin/env bash
# Round 2, GPU call 18: basis stores by the gather warps (C2_IO_OUT)
set -u
mkdir -p gpurun_out
for name in ioout ioout56; do
  V=$PWD/deepsphere-cosmo-tf2_b200/lib/libdeepsphere_b200_$name.so
  [ -f "$V" ] || continue
  {
    DEEPSPHERE_LIB=$V timeout 400 python -m pytest tests/test_gpu_lattice.py tests/test_gpu_r2_shapes.py -q -m gpu -k "fused or conv2 or tf32" 2>&1 | tail -12
    DEEPSPHERE_LIB=$V timeout 200 python tools/bench_fwd.py tf32 32 2>&1 | grep RESULT
    DEEPSPHERE_LIB=$V timeout 200 python tools/bench_fwd.py tf32 32 bwd 2>&1 | grep RESULT
  } > gpurun_out/r2r_variant_$name.log 2>&1
  echo "$name: $(grep -E 'passed|failed' gpurun_out/r2r_variant_$name.log | tail -1) | $(grep RESULT gpurun_out/r2r_variant_$name.log | cut -c1-100 | tr '\n' '|')"
  grep -m2 -E "Error|error" gpurun_out/r2r_variant_$name.log | cut -c1-300
done
V=$PWD/deepsphere-cosmo-tf2_b200/lib/libdeepsphere_b200_ioout.so
DEEPSPHERE_LIB=$V timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-other-modes --no-configs --no-f-sweep 2>/dev/null | grep '^{' > gpurun_out/r2r_bench_ioout.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2r_bench_ioout.json'))
print('ioout layer', round(d['ms_per_step'], 2), 'fwd', round(d['kernels']['forward']['ms'], 2), 'bwd', round(d['kernels']['backward']['ms'], 2), 'model_train', round(d['model_train']['ms_per_step'], 3), 'partitioned', round(d['model_train_partitioned']['ms_per_step'], 2))
PY
