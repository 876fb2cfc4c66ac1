// tcgen05 tensor-core contraction — placeholder until the UMMA kernels land (next commit).
#include "ds_common.cuh"
namespace ds {
int umma_supported(int64_t, int, int64_t) { return -1; }
int launch_umma_gemm_nn(int64_t, int64_t, int64_t, int, const float*, const float*, int64_t, const float*, int64_t,
                        int64_t, const float*, int, float*, int, cudaStream_t) {
  return fail("tensor-core mode not built");
}
}  // namespace ds
