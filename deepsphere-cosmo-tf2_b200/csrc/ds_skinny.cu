// Streaming kernels for "skinny" contractions: the pseudo-convolutions at the head of a HealpyGCNN
// (healpy_layers.py:87-146, Conv1D with kernel = stride = 4^p over NESTED-contiguous children) have a reduction
// length of only Kc = 4^p * Fin (4 for the usual p = 1 on a single-channel map) and a few output channels.  The
// 64x64x16-tiled fp32 GEMMs of ds_gemm.cu spend such a problem on padding (0.77 ms forward + 0.76 ms weight gradient
// at nside 256, batch 16, 1 -> 16 channels: 20x the time the 250 MB of traffic need); these kernels stream it: one
// thread per (output row, 4-channel group), the tiny weight matrix read through L1, the backward pass fused into ONE
// sweep over x, y, dy (activation derivative, weight gradient, bias gradient; dz only written when dx is wanted).
//
// HBM-bound by construction: forward reads R*Kc and writes R*N floats; backward reads R*(Kc + 2N) floats.
//
// Round 1 wrote these after its GPU budget was spent (host-emulated only, opt-in).  Round 2 measured them on the B200 -
// HealpyGCNN training step nside 256 / batch 16: 7.94 -> 5.85 ms; the sphere-partitioned nside-1024 step: 55.9 -> 39.5 ms
// (the first pseudo-convolution alone was 30 % of it), same losses and gradients - and made them the default;
// DEEPSPHERE_SKINNY=0 switches back to the tiled kernels.
#include <algorithm>
#include <cstdlib>

#include "ds_common.cuh"

namespace ds {
namespace {

constexpr int SK_THREADS = 256;
constexpr int SK_MAX_N = 64;

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// C[r, :] = act( A[r, :KC] * Bm[KC, N] + bias ),  thread = (row r, column group g of 4)
template <int KC>
__global__ void __launch_bounds__(SK_THREADS) skinny_nn_kernel(int64_t R, int N, const float* __restrict__ A,
                                                               const float* __restrict__ Bm,
                                                               const float* __restrict__ bias, int bias_mod, int act,
                                                               float* __restrict__ C) {
  const int cg = N >> 2;
  const int64_t total = R * cg;
  for (int64_t e = (int64_t)blockIdx.x * SK_THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * SK_THREADS) {
    const int64_t r = e / cg;
    const int g = (int)(e - r * cg);
    float a[KC];
#pragma unroll
    for (int k4 = 0; k4 < KC / 4; ++k4) {
      const float4 v = ldg4(A + r * KC + 4 * k4);
      a[4 * k4] = v.x; a[4 * k4 + 1] = v.y; a[4 * k4 + 2] = v.z; a[4 * k4 + 3] = v.w;
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < KC; ++k) {  // ascending k, like the tiled kernel: same rounding
      const float4 b = ldg4(Bm + (int64_t)k * N + 4 * g);
      acc.x = fmaf(a[k], b.x, acc.x); acc.y = fmaf(a[k], b.y, acc.y);
      acc.z = fmaf(a[k], b.z, acc.z); acc.w = fmaf(a[k], b.w, acc.w);
    }
    if (bias != nullptr) {
      const int n = 4 * g;
      acc.x += __ldg(bias + (n % bias_mod)); acc.y += __ldg(bias + ((n + 1) % bias_mod));
      acc.z += __ldg(bias + ((n + 2) % bias_mod)); acc.w += __ldg(bias + ((n + 3) % bias_mod));
    }
    if (act != DS_ACT_LINEAR) {
      acc.x = act_apply(acc.x, act); acc.y = act_apply(acc.y, act);
      acc.z = act_apply(acc.z, act); acc.w = act_apply(acc.w, act);
    }
    reinterpret_cast<float4*>(C + r * N)[g] = acc;
  }
}

// One sweep of the backward pass.  Thread = (row slot, column group g); g is fixed per thread (cg divides 32), the
// rows advance by the number of slots.  Per thread: dz = dy * act'(y); acc[k] += x[r, k] * dz (k < KC);
// acc[KC] += dz.  Then the lanes with equal g are summed with shuffles, the 8 warps through shared memory, and the
// block writes partial[block][KC + 1][N] (fixed order everywhere: deterministic).
template <int KC>
__global__ void __launch_bounds__(SK_THREADS) skinny_bwd_kernel(int64_t R, int N, const float* __restrict__ X,
                                                                const float* __restrict__ y,
                                                                const float* __restrict__ dy, int act,
                                                                float* __restrict__ dz_out,
                                                                float* __restrict__ partial) {
  __shared__ __align__(16) float red[SK_THREADS / 32][SK_MAX_N];
  const int cg = N >> 2;  // 1, 2, 4, 8 or 16
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane % cg;
  const int64_t slots = ((int64_t)gridDim.x * SK_THREADS) / cg;
  const int64_t slot = ((int64_t)blockIdx.x * SK_THREADS + tid) / cg;
  float4 acc[KC + 1];
#pragma unroll
  for (int k = 0; k <= KC; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t r = slot; r < R; r += slots) {
    float4 d = ldg4(dy + r * N + 4 * g);
    if (act != DS_ACT_LINEAR) {
      const float4 yv = ldg4(y + r * N + 4 * g);
      d.x *= act_grad_from_y(yv.x, act); d.y *= act_grad_from_y(yv.y, act);
      d.z *= act_grad_from_y(yv.z, act); d.w *= act_grad_from_y(yv.w, act);
      if (dz_out != nullptr) reinterpret_cast<float4*>(dz_out + r * N)[g] = d;
    }
#pragma unroll
    for (int k4 = 0; k4 < KC / 4; ++k4) {
      const float4 v = ldg4(X + r * KC + 4 * k4);
      const float a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4& c = acc[4 * k4 + j];
        c.x = fmaf(a[j], d.x, c.x); c.y = fmaf(a[j], d.y, c.y);
        c.z = fmaf(a[j], d.z, c.z); c.w = fmaf(a[j], d.w, c.w);
      }
    }
    acc[KC].x += d.x; acc[KC].y += d.y; acc[KC].z += d.z; acc[KC].w += d.w;
  }
#pragma unroll
  for (int k = 0; k <= KC; ++k) {
    float4 v = acc[k];
    for (int off = cg; off < 32; off <<= 1) {  // lanes g, g + cg, g + 2 cg, ... hold the same column group
      v.x += __shfl_xor_sync(0xffffffffu, v.x, off);
      v.y += __shfl_xor_sync(0xffffffffu, v.y, off);
      v.z += __shfl_xor_sync(0xffffffffu, v.z, off);
      v.w += __shfl_xor_sync(0xffffffffu, v.w, off);
    }
    if (lane < cg) *reinterpret_cast<float4*>(&red[warp][4 * lane]) = v;
    __syncthreads();
    if (tid < N) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < SK_THREADS / 32; ++w) s += red[w][tid];
      partial[((int64_t)blockIdx.x * (KC + 1) + k) * N + tid] = s;
    }
    __syncthreads();
  }
}

// dw[k, n] (k < KC) and dbias[n] = sum over the blocks' partials
__global__ void skinny_bwd_final_kernel(int N, int KC, int nblk, const float* __restrict__ partial,
                                        float* __restrict__ dw, float* __restrict__ dbias) {
  const int total = (KC + 1) * N;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < nblk; ++b) s += __ldg(partial + (int64_t)b * total + e);
    if (e < KC * N) dw[e] = s;
    else if (dbias != nullptr) dbias[e - KC * N] = s;
  }
}

// dz = dy * act'(y) and the column sums of dz in ONE sweep (act_backward_kernel + colsum_partial_kernel read dz a second
// time).  Thread = (row slot, column group of 4); a block owns a contiguous row range; partial[block][F] feeds the
// existing deterministic final reduction (colsum_final_kernel's layout).  F / 4 must divide 256.
__global__ void __launch_bounds__(SK_THREADS) act_backward_colsum_kernel(int64_t R, int F, const float* __restrict__ y,
                                                                         const float* __restrict__ dy, int act,
                                                                         float* __restrict__ dz,
                                                                         float* __restrict__ partial) {
  __shared__ __align__(16) float4 red4[SK_THREADS];
  const int cg = F >> 2, tid = threadIdx.x;
  const int g = tid % cg, slot = tid / cg, slots = SK_THREADS / cg;
  const int64_t rows_per_block = (R + gridDim.x - 1) / gridDim.x;
  const int64_t rb = (int64_t)blockIdx.x * rows_per_block;
  const int64_t re = rb + rows_per_block < R ? rb + rows_per_block : R;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t r = rb + slot; r < re; r += slots) {
    float4 d = ldg4(dy + r * F + 4 * g);
    const float4 yv = ldg4(y + r * F + 4 * g);
    d.x *= act_grad_from_y(yv.x, act); d.y *= act_grad_from_y(yv.y, act);
    d.z *= act_grad_from_y(yv.z, act); d.w *= act_grad_from_y(yv.w, act);
    reinterpret_cast<float4*>(dz + r * F)[g] = d;
    acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w;
  }
  red4[tid] = acc;
  __syncthreads();
  if (tid < F) {
    const int gg = tid >> 2, j = tid & 3;
    float s = 0.f;
    for (int q = 0; q < slots; ++q) {
      const float4 v = red4[q * cg + gg];
      s += j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w));
    }
    partial[(int64_t)blockIdx.x * F + tid] = s;
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline bool enabled() {  // on by default since round 2 (DEEPSPHERE_SKINNY=0 falls back to the tiled kernels)
  static const bool on = [] { const char* e = getenv("DEEPSPHERE_SKINNY"); return e == nullptr || atoi(e) != 0; }();
  return on;
}

inline bool shape_ok(int64_t Kc, int64_t N) {
  if (!(Kc == 4 || Kc == 8 || Kc == 12 || Kc == 16)) return false;
  return N == 4 || N == 8 || N == 16 || N == 32 || N == 64;
}

inline int bwd_blocks(int64_t R, int64_t N) {
  const int64_t threads = R * (N / 4);
  return (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)2 * num_sms(), (threads + SK_THREADS - 1) / SK_THREADS));
}

}  // namespace

#ifndef DS_EMULATE  // tests/emul compiles the kernels above for the host; the launchers below need nvcc

bool skinny_usable(int64_t R, int64_t Kc, int64_t N) { return enabled() && R >= 1 && shape_ok(Kc, N); }

int64_t skinny_bwd_workspace_elems(int64_t R, int64_t Kc, int64_t N) {
  return shape_ok(Kc, N) ? (int64_t)bwd_blocks(R, N) * (Kc + 1) * N : 0;
}

// C[R, N] = act( A[R, Kc] * Bm[Kc, N] + bias[n % bias_mod] ), all operands dense row-major.  Returns -1 (nothing
// launched) when the pointers are not 16-byte aligned: the caller falls through to the tiled kernel; 0 on success.
int launch_skinny_nn(int64_t R, int64_t N, int64_t Kc, const float* A, const float* Bm, const float* bias,
                     int64_t bias_mod, int act, float* C, cudaStream_t st) {
  if (!aligned16(A) || !aligned16(Bm) || !aligned16(C)) return -1;
  const int64_t total = R * (N / 4);
  const unsigned blocks = (unsigned)std::min<int64_t>((total + SK_THREADS - 1) / SK_THREADS, (int64_t)num_sms() * 16);
  const int bm = (int)(bias_mod > 0 ? bias_mod : 1);
#define DS_SK_NN(KC) \
  skinny_nn_kernel<KC><<<blocks, SK_THREADS, 0, st>>>(R, (int)N, A, Bm, bias, bm, act, C)
  switch (Kc) {
    case 4: DS_SK_NN(4); break;
    case 8: DS_SK_NN(8); break;
    case 12: DS_SK_NN(12); break;
    default: DS_SK_NN(16); break;
  }
#undef DS_SK_NN
  DS_LAUNCHED();
  return 0;
}

// dw[Kc, N] = X^T dz, dbias[N] = colsum(dz) with dz = dy * act'(y); dz_out (optional, only used when act != LINEAR)
// receives dz.  partial: skinny_bwd_workspace_elems floats.  Returns -1 on misalignment (nothing launched).
int launch_skinny_pconv_bwd(int64_t R, int64_t N, int64_t Kc, const float* X, const float* y, const float* dy, int act,
                            float* dz_out, float* dw, float* dbias, float* partial, cudaStream_t st) {
  if (!aligned16(X) || !aligned16(dy) || (act != DS_ACT_LINEAR && (!aligned16(y) || (dz_out && !aligned16(dz_out)))))
    return -1;
  const int nblk = bwd_blocks(R, N);
#define DS_SK_BWD(KC) \
  skinny_bwd_kernel<KC><<<nblk, SK_THREADS, 0, st>>>(R, (int)N, X, y, dy, act, dz_out, partial)
  switch (Kc) {
    case 4: DS_SK_BWD(4); break;
    case 8: DS_SK_BWD(8); break;
    case 12: DS_SK_BWD(12); break;
    default: DS_SK_BWD(16); break;
  }
#undef DS_SK_BWD
  DS_LAUNCHED();
  const int total = (int)((Kc + 1) * N);
  skinny_bwd_final_kernel<<<(total + 127) / 128, 128, 0, st>>>((int)N, (int)Kc, nblk, partial, dw, dbias);
  DS_LAUNCHED();
  return 0;
}

int launch_act_backward_colsum(int64_t R, int64_t F, const float* y, const float* dy, int act, float* dz, float* dbias,
                               float* workspace, cudaStream_t st) {
  if (!enabled() || act == DS_ACT_LINEAR || dbias == nullptr || R < 1) return -1;
  if (F < 4 || F > SK_THREADS || (F & (F - 1)) != 0) return -1;  // F/4 must divide 256 and F <= 256 threads
  if (!aligned16(y) || !aligned16(dy) || !aligned16(dz)) return -1;
  const int64_t max_blocks = colsum_workspace_elems(F) / F;
  const int nblk = (int)std::max<int64_t>(1, std::min<int64_t>(max_blocks, (R + 63) / 64));
  act_backward_colsum_kernel<<<nblk, SK_THREADS, 0, st>>>(R, (int)F, y, dy, act, dz, workspace);
  DS_LAUNCHED();
  return launch_colsum_final(F, F, nblk, workspace, dbias, st);
}

#endif  // DS_EMULATE

}  // namespace ds
