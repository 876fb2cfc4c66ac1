#!/usr/bin/env bash
# Round 2, GPU call 3: looped hop bodies (instruction-cache footprint) + the round's new parity tests
set -u
mkdir -p gpurun_out
for name in ${VARIANTS:-default loop loopsymw bc6loop bc6 noprobe}; do
  V=$PWD/deepsphere-cosmo-tf2_b200/lib/libdeepsphere_b200_$name.so
  [ "$name" = default ] && V=$PWD/deepsphere-cosmo-tf2_b200/lib/libdeepsphere_b200.so
  [ -f "$V" ] || continue
  {
    echo "== $name: parity of the fused kernel =="
    DEEPSPHERE_LIB=$V timeout 300 python -m pytest tests/test_gpu_lattice.py -q -m gpu -k "fused or conv2" 2>&1 | tail -4
    echo "== $name: fwd / fwd+bwd =="
    DEEPSPHERE_LIB=$V timeout 200 python tools/bench_fwd.py tf32 32 2>&1 | grep RESULT
    DEEPSPHERE_LIB=$V timeout 200 python tools/bench_fwd.py tf32 32 bwd 2>&1 | grep RESULT
    DEEPSPHERE_LIB=$V timeout 300 ncu --metrics sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:lattice_conv2 -c 1 python tools/profile_layer.py tf32 8 2>&1 | grep -E "icc|gcc|duration|issue_active|pipe_fma"
  } > gpurun_out/r2c_variant_$name.log 2>&1
  echo "$name: $(grep -E 'passed|failed|rror' gpurun_out/r2c_variant_$name.log | tail -1) | $(grep RESULT gpurun_out/r2c_variant_$name.log | cut -c1-100 | tr '\n' '|')"
  grep -E "icc|gcc|duration|issue_active|pipe_fma" gpurun_out/r2c_variant_$name.log | awk '{print $1, $NF}' | tr '\n' ';'; echo
done
timeout 900 python -m pytest tests/test_gpu_r2_shapes.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2c_tests_shapes.log
tail -5 gpurun_out/r2c_tests_shapes.log
