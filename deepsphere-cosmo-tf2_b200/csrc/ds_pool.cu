// HealpyPool (healpy_layers.py:20-84) and the bias+activation epilogue that follows an
// optional BatchNorm (gnn_layers.py:155-159).  Pure streaming kernels: in NESTED order the
// 4^p children of a coarse pixel are contiguous rows, so pooling is a reshape-reduce.
// HBM-bound: forward moves (1 + 4^-p) * B*M*F*4 bytes.
#include "ds_common.cuh"

namespace ds {
namespace {

// one thread per output vector (V channels of one coarse pixel)
template <int V>
__global__ void __launch_bounds__(256) pool_fwd_kernel(int64_t n_out_vec, int64_t FV, int r, int type,
                                                       const float* __restrict__ x, float* __restrict__ y) {
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n_out_vec; o += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = o / FV, fv = o - row * FV;
    const float* src = x + (row * r * FV + fv) * V;
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = type == DS_POOL_MAX ? -INFINITY : 0.f;
    for (int c = 0; c < r; ++c) {
      float v[V];
      if (V == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(src + (int64_t)c * FV * V));
        v[0] = t.x; v[1 % V] = t.y; v[2 % V] = t.z; v[3 % V] = t.w;
      } else {
        v[0] = __ldg(src + (int64_t)c * FV * V);
      }
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] = type == DS_POOL_MAX ? fmaxf(acc[i], v[i]) : acc[i] + v[i];
    }
    if (type == DS_POOL_AVG) {
      const float inv = 1.f / (float)r;
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] *= inv;
    }
#pragma unroll
    for (int i = 0; i < V; ++i) y[o * V + i] = acc[i];
  }
}

// one thread per output element: recompute the arg-max (first maximum wins, as TF's
// MaxPoolGrad) and route dy to it; AVG spreads dy / 4^p.
__global__ void __launch_bounds__(256) pool_bwd_kernel(int64_t n_out, int64_t F, int r, int type,
                                                       const float* __restrict__ x, const float* __restrict__ dy,
                                                       float* __restrict__ dx) {
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n_out; o += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = o / F, f = o - row * F;
    const int64_t base = row * r * F + f;
    const float g = __ldg(dy + o);
    if (type == DS_POOL_AVG) {
      const float v = g / (float)r;
      for (int c = 0; c < r; ++c) dx[base + (int64_t)c * F] = v;
    } else {
      int arg = 0;
      float best = __ldg(x + base);
      for (int c = 1; c < r; ++c) {
        const float v = __ldg(x + base + (int64_t)c * F);
        if (v > best) { best = v; arg = c; }
      }
      for (int c = 0; c < r; ++c) dx[base + (int64_t)c * F] = c == arg ? g : 0.f;
    }
  }
}

__global__ void __launch_bounds__(256) bias_act_kernel(int64_t n, int64_t F, const float* __restrict__ z,
                                                       const float* __restrict__ bias, int act, float* __restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = z[i];
    if (bias != nullptr) v += __ldg(bias + (i % F));
    y[i] = act_apply(v, act);
  }
}

inline unsigned grid_for(int64_t n) {
  const int64_t b = (n + 255) / 256;
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(b, (int64_t)num_sms() * 32));
}

int pool_check(int64_t B, int64_t M, int64_t F, int32_t p, int32_t type, const char* who) {
  DS_CHECK(p >= 1 && p <= 12, "%s: p=%d out of range (healpy_layers.py:39-40 requires p >= 1)", who, p);
  DS_CHECK(type == DS_POOL_MAX || type == DS_POOL_AVG, "%s: pooling type %d not understood", who, type);
  DS_CHECK(B >= 1 && M >= 1 && F >= 1, "%s: empty shape", who);
  DS_CHECK(M % (1LL << (2 * p)) == 0, "%s: M=%lld not compatible with the filter size %lld", who, (long long)M,
           (long long)(1LL << (2 * p)));
  return 0;
}

}  // namespace
}  // namespace ds

extern "C" {

int ds_pool_forward(int64_t B, int64_t M, int64_t F, int32_t p, int32_t pool_type, const float* x, float* y,
                    void* stream) {
  using namespace ds;
  DS_TRY(pool_check(B, M, F, p, pool_type, "ds_pool_forward"));
  DS_CHECK(x && y, "ds_pool_forward: NULL tensor");
  const int r = 1 << (2 * p);
  const int64_t n_out = B * (M / r) * F;
  const bool vec4 = (F % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  if (vec4) pool_fwd_kernel<4><<<grid_for(n_out / 4), 256, 0, (cudaStream_t)stream>>>(n_out / 4, F / 4, r, pool_type, x, y);
  else pool_fwd_kernel<1><<<grid_for(n_out), 256, 0, (cudaStream_t)stream>>>(n_out, F, r, pool_type, x, y);
  DS_LAUNCHED();
  return 0;
}

int ds_pool_backward(int64_t B, int64_t M, int64_t F, int32_t p, int32_t pool_type, const float* x, const float* dy,
                     float* dx, void* stream) {
  using namespace ds;
  DS_TRY(pool_check(B, M, F, p, pool_type, "ds_pool_backward"));
  DS_CHECK(dy && dx && (pool_type == DS_POOL_AVG || x), "ds_pool_backward: NULL tensor");
  const int r = 1 << (2 * p);
  const int64_t n_out = B * (M / r) * F;
  pool_bwd_kernel<<<grid_for(n_out), 256, 0, (cudaStream_t)stream>>>(n_out, F, r, pool_type, x, dy, dx);
  DS_LAUNCHED();
  return 0;
}

int ds_bias_act_forward(int64_t R, int64_t F, const float* z, const float* bias, int32_t act, float* y, void* stream) {
  using namespace ds;
  DS_CHECK(z && y && R > 0 && F > 0, "ds_bias_act_forward: bad argument");
  DS_CHECK(act >= DS_ACT_LINEAR && act <= DS_ACT_SOFTPLUS, "ds_bias_act_forward: unknown activation id %d", act);
  bias_act_kernel<<<grid_for(R * F), 256, 0, (cudaStream_t)stream>>>(R * F, F, z, bias, act, y);
  DS_LAUNCHED();
  return 0;
}

}  // extern "C"
