#!/usr/bin/env bash
# First GPU call of the next round: everything that was written after round 1's GPU budget ran out gets measured here.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
# Outputs land in gpurun_out/ (merged back by gpurun).
set -u
mkdir -p gpurun_out
{
  echo "== 1. GPU tests of the rows added without a GPU (Bernstein, HealpySmoothing, streaming kernels) =="
  python -m pytest tests/test_gpu_zz_next.py -q -m gpu 2>&1 | tail -15
  echo "== 2. the whole GPU suite =="
  python -m pytest tests -x -q -m gpu 2>&1 | tail -5
  echo "== 3. smoke =="
  python __graft_entry__.py --smoke 2>&1 | tail -2
} > gpurun_out/r2_tests.log 2>&1
# 4. default bench line (includes model_train_experimental = the HealpyGCNN step with DEEPSPHERE_SKINNY=1, validated)
python bench.py > gpurun_out/r2_bench_default.log 2>&1
# 5. launch list of the model step with and without the streaming kernels (kernel shares; ncu times are cold-cache)
for sk in 0 1; do
  DEEPSPHERE_SKINNY=$sk ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/r2_launches_model_skinny$sk.csv python tools/profile_model.py > gpurun_out/r2_model_prof_skinny$sk.log 2>&1
done
# 6. fused-kernel experiments (build them HERE first: bash tools/build_variants.sh; the .so files travel with the
#    snapshot).  Every variant has run on the host emulator (tests/test_emul_cpu.py); this decides parity and speed.
# VARIANTS="fence all" bash tools/round2_first_call.sh   restricts the list (about 2 GPU-minutes per variant)
for name in ${VARIANTS:-split fence noprobe epipipe symw br3all br2r144 br2all br2allfence}; do
  V=deepsphere-cosmo-tf2_b200/lib/libdeepsphere_b200_$name.so
  [ -f "$V" ] || continue
  {
    echo "== $name: parity of the fused kernel =="
    DEEPSPHERE_LIB=$PWD/$V timeout 600 python -m pytest tests/test_gpu_lattice.py -q -m gpu -k "fused or conv2" 2>&1 | tail -4
    echo "== $name: layer bench =="
    DEEPSPHERE_LIB=$PWD/$V timeout 600 python bench.py --steps 5 --warmup 3 --no-model --no-e2e --no-cpu-baseline --no-other-modes 2>&1 | tail -1 | cut -c1-1400
  } > gpurun_out/r2_variant_$name.log 2>&1
  tail -c 500 gpurun_out/r2_variant_$name.log
done
tail -c 600 gpurun_out/r2_tests.log
tail -c 1500 gpurun_out/r2_bench_default.log
