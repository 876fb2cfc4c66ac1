// Halo pack / assemble / reduce for the sphere-partitioned path (SURVEY 8e.2, 8b "ds_halo_*").
//
// The reference is single-device; a partitioned graph convolution needs, per layer, the (K-1)-hop neighbourhood of a
// rank's own rows.  The exchange itself is one all-to-all (NCCL, issued by the host framework); what runs on the GPU
// either side of it is pure row gather / scatter in the native [B, rows, F] layout.  Three kernels replace the
// index_select / permute / cat / index_copy_ / index_add_ chains (4 - 6 framework kernels per peer and direction):
//   ds_halo_pack      send[j, b, :] = x[b, rows[j], :]            all peers' rows in ONE launch (peer blocks contiguous)
//   ds_halo_assemble  x_ext = own rows (copied as one slab per sample) + received halo rows at their positions
//   ds_halo_reduce    g_own = g_ext[own slab] + sum of the contributions the peers returned, in a FIXED order per
//                     row (a CSR list row -> receive slots built on the host): bit-reproducible, no atomics
// HBM-bound streaming kernels; rows are F contiguous floats (float4 granularity when F % 4 == 0).
#include <algorithm>

#include "ds_common.cuh"

namespace ds {
namespace {

inline unsigned halo_grid(int64_t n) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 16));
}

// dst[j, b, v] = src[b, rows[j], v]      (V floats per element)
template <typename T>
__global__ void __launch_bounds__(256) halo_pack_kernel(int64_t B, int64_t n_src, int64_t n, int64_t FV,
                                                        const int32_t* __restrict__ rows, const T* __restrict__ src,
                                                        T* __restrict__ dst) {
  const int64_t total = n * B * FV;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = e % FV, b = (e / FV) % B, j = e / (FV * B);
    dst[e] = __ldg(src + (b * n_src + rows[j]) * FV + v);
  }
}

// x_ext[b, own_start + r, :] = x_own[b, r, :];  x_ext[b, pos[j], :] = recv[j, b, :]
template <typename T>
__global__ void __launch_bounds__(256) halo_assemble_kernel(int64_t B, int64_t n_own, int64_t n_ext, int64_t own_start,
                                                            int64_t n_halo, int64_t FV, const int32_t* __restrict__ pos,
                                                            const T* __restrict__ x_own, const T* __restrict__ recv,
                                                            T* __restrict__ x_ext) {
  const int64_t n_copy = B * n_own * FV, total = n_copy + n_halo * B * FV;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    if (e < n_copy) {
      const int64_t v = e % FV, r = (e / FV) % n_own, b = e / (FV * n_own);
      x_ext[(b * n_ext + own_start + r) * FV + v] = __ldg(x_own + e);
    } else {
      const int64_t h = e - n_copy;
      const int64_t v = h % FV, b = (h / FV) % B, j = h / (FV * B);
      x_ext[(b * n_ext + pos[j]) * FV + v] = __ldg(recv + h);
    }
  }
}

template <typename T>
__device__ __forceinline__ T halo_add(T a, T b);
template <>
__device__ __forceinline__ float halo_add<float>(float a, float b) { return a + b; }
template <>
__device__ __forceinline__ float4 halo_add<float4>(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// g_own[b, r, :] = g_ext[b, own_start + r, :]                                   (every own row)
// g_own[b, urow[u], :] += sum_{s in slots[uptr[u] .. uptr[u+1])} recv[s, b, :]   (rows some peer returned, fixed order)
template <typename T>
__global__ void __launch_bounds__(256) halo_reduce_kernel(int64_t B, int64_t n_own, int64_t n_ext, int64_t own_start,
                                                          int64_t FV, const int32_t* __restrict__ row_slot_ptr,
                                                          const int32_t* __restrict__ slots, const T* __restrict__ g_ext,
                                                          const T* __restrict__ recv, T* __restrict__ g_own) {
  const int64_t total = B * n_own * FV;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = e % FV, r = (e / FV) % n_own, b = e / (FV * n_own);
    T acc = __ldg(g_ext + (b * n_ext + own_start + r) * FV + v);
    if (row_slot_ptr != nullptr) {
      const int32_t s0 = row_slot_ptr[r], s1 = row_slot_ptr[r + 1];
      for (int32_t s = s0; s < s1; ++s) acc = halo_add(acc, __ldg(recv + ((int64_t)slots[s] * B + b) * FV + v));
    }
    g_own[e] = acc;
  }
}

}  // namespace
}  // namespace ds

extern "C" int ds_halo_pack(int64_t B, int64_t n_src, int64_t F, int64_t n, const int32_t* rows, const float* src,
                            float* dst, void* stream) {
  using namespace ds;
  DS_CHECK(B >= 0 && n_src >= 0 && F >= 1 && n >= 0, "ds_halo_pack: bad sizes");
  if (n == 0 || B == 0) return 0;
  DS_CHECK(rows && src && dst, "ds_halo_pack: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (F % 4 == 0)
    halo_pack_kernel<float4><<<halo_grid(n * B * (F / 4)), 256, 0, st>>>(B, n_src, n, F / 4, rows,
                                                                          reinterpret_cast<const float4*>(src),
                                                                          reinterpret_cast<float4*>(dst));
  else
    halo_pack_kernel<float><<<halo_grid(n * B * F), 256, 0, st>>>(B, n_src, n, F, rows, src, dst);
  DS_LAUNCHED();
  return 0;
}

extern "C" int ds_halo_assemble(int64_t B, int64_t n_own, int64_t n_ext, int64_t own_start, int64_t F, int64_t n_halo,
                                const int32_t* pos, const float* x_own, const float* recv, float* x_ext, void* stream) {
  using namespace ds;
  DS_CHECK(B >= 0 && n_own >= 0 && n_ext >= n_own && own_start >= 0 && own_start + n_own <= n_ext && F >= 1 && n_halo >= 0,
           "ds_halo_assemble: bad sizes");
  if (B == 0 || n_ext == 0) return 0;
  DS_CHECK(x_own && x_ext && (n_halo == 0 || (pos && recv)), "ds_halo_assemble: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (F % 4 == 0)
    halo_assemble_kernel<float4><<<halo_grid((B * n_own + n_halo * B) * (F / 4)), 256, 0, st>>>(
        B, n_own, n_ext, own_start, n_halo, F / 4, pos, reinterpret_cast<const float4*>(x_own),
        reinterpret_cast<const float4*>(recv), reinterpret_cast<float4*>(x_ext));
  else
    halo_assemble_kernel<float><<<halo_grid((B * n_own + n_halo * B) * F), 256, 0, st>>>(B, n_own, n_ext, own_start, n_halo,
                                                                                       F, pos, x_own, recv, x_ext);
  DS_LAUNCHED();
  return 0;
}

extern "C" int ds_halo_reduce(int64_t B, int64_t n_own, int64_t n_ext, int64_t own_start, int64_t F,
                              const int32_t* row_slot_ptr, const int32_t* slots, const float* g_ext, const float* recv,
                              float* g_own, void* stream) {
  using namespace ds;
  DS_CHECK(B >= 0 && n_own >= 0 && n_ext >= n_own && own_start >= 0 && own_start + n_own <= n_ext && F >= 1,
           "ds_halo_reduce: bad sizes");
  if (B == 0 || n_own == 0) return 0;
  DS_CHECK(g_ext && g_own && (row_slot_ptr == nullptr || (slots && recv)), "ds_halo_reduce: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (F % 4 == 0)
    halo_reduce_kernel<float4><<<halo_grid(B * n_own * (F / 4)), 256, 0, st>>>(
        B, n_own, n_ext, own_start, F / 4, row_slot_ptr, slots, reinterpret_cast<const float4*>(g_ext),
        reinterpret_cast<const float4*>(recv), reinterpret_cast<float4*>(g_own));
  else
    halo_reduce_kernel<float><<<halo_grid(B * n_own * F), 256, 0, st>>>(B, n_own, n_ext, own_start, F, row_slot_ptr, slots,
                                                                       g_ext, recv, g_own);
  DS_LAUNCHED();
  return 0;
}
