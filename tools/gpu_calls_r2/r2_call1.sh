#!/usr/bin/env bash
# Round 2, GPU call 1: whole GPU suite + default bench (r2 baseline) + the prepared fused-kernel variants.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_gpu.txt 2>&1
{
  echo "== whole GPU suite =="
  timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
  echo "== smoke =="
  timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
} > gpurun_out/r2_tests.log 2>&1
timeout 900 python bench.py > gpurun_out/r2_bench_default.log 2>&1
for name in ${VARIANTS:-split noprobe epipipe symw br3all br2all br2r144 fence r216}; do
  V=deepsphere-cosmo-tf2_b200/lib/libdeepsphere_b200_$name.so
  [ -f "$V" ] || continue
  {
    echo "== $name: parity of the fused kernel =="
    DEEPSPHERE_LIB=$PWD/$V timeout 300 python -m pytest tests/test_gpu_lattice.py -q -m gpu -k "fused or conv2" 2>&1 | tail -4
    echo "== $name: layer bench =="
    DEEPSPHERE_LIB=$PWD/$V timeout 300 python bench.py --steps 5 --warmup 3 --no-model --no-e2e --no-cpu-baseline --no-other-modes 2>&1 | tail -1 | cut -c1-1600
  } > gpurun_out/r2_variant_$name.log 2>&1
  echo "$name: $(grep -o '"fused_forward": {[^}]*}' gpurun_out/r2_variant_$name.log | grep -o '"ms": [0-9.]*') $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_variant_$name.log | head -1) $(grep -E 'passed|failed' gpurun_out/r2_variant_$name.log | tail -1)"
done
tail -c 400 gpurun_out/r2_tests.log
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_bench_default.log | head -1
