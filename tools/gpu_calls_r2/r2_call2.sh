#!/usr/bin/env bash
# Round 2, GPU call 2: the new block shapes of the fused kernel (3x6 pixels x 2 channels; 3x3 x 2 channels with 8 warps)
set -u
mkdir -p gpurun_out
for name in ${VARIANTS:-bc6 bc6io40 bc6split bc6noskip cpt2 cpt2n32}; do
  V=$PWD/deepsphere-cosmo-tf2_b200/lib/libdeepsphere_b200_$name.so
  [ -f "$V" ] || continue
  {
    echo "== $name: parity of the fused kernel =="
    DEEPSPHERE_LIB=$V timeout 300 python -m pytest tests/test_gpu_lattice.py -q -m gpu -k "fused or conv2" 2>&1 | tail -4
    echo "== $name: fwd / fwd+bwd =="
    DEEPSPHERE_LIB=$V timeout 200 python tools/bench_fwd.py tf32 32 2>&1 | grep RESULT
    DEEPSPHERE_LIB=$V timeout 200 python tools/bench_fwd.py tf32 32 bwd 2>&1 | grep RESULT
  } > gpurun_out/r2b_variant_$name.log 2>&1
  echo "$name: $(grep -E 'passed|failed|rror' gpurun_out/r2b_variant_$name.log | tail -1) | $(grep RESULT gpurun_out/r2b_variant_$name.log | cut -c1-110 | tr '\n' '|')"
done
# one full ncu capture of the 3x6x2 kernel (batch 4: kernel replay stays cheap)
V=$PWD/deepsphere-cosmo-tf2_b200/lib/libdeepsphere_b200_bc6.so
DEEPSPHERE_LIB=$V timeout 600 ncu --set full --import-source on --clock-control none -k regex:lattice_conv2 -c 1 -f -o gpurun_out/r2b_prof_bc6 python tools/profile_layer.py tf32 8 > gpurun_out/r2b_ncu.log 2>&1
tail -3 gpurun_out/r2b_ncu.log
