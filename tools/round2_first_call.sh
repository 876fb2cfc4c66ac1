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
# 6. fused-kernel experiment: proxy fence by the issuing lane (build it HERE first, the .so travels with the snapshot:
#    python deepsphere-cosmo-tf2_b200/build.py --variant fence -DC2_FENCE_BY_ISSUER=1)
V=deepsphere-cosmo-tf2_b200/lib/libdeepsphere_b200_fence.so
if [ -f "$V" ]; then
  {
    echo "== fence variant: parity of the fused kernel =="
    DEEPSPHERE_LIB=$PWD/$V python -m pytest tests/test_gpu_lattice.py tests/test_gpu_tensor_core.py -q -m gpu 2>&1 | tail -5
    echo "== fence variant: layer bench =="
    DEEPSPHERE_LIB=$PWD/$V python bench.py --no-model --no-e2e --no-cpu-baseline --no-other-modes 2>&1 | tail -1
  } > gpurun_out/r2_variant_fence.log 2>&1
  tail -c 900 gpurun_out/r2_variant_fence.log
fi
tail -c 600 gpurun_out/r2_tests.log
tail -c 1500 gpurun_out/r2_bench_default.log
