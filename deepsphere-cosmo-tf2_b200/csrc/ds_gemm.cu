// DS_MODE_FP32 contraction kernels: exact-fp32 CUDA-core GEMMs for the K*Fin -> Fout
// feature contraction (gnn_layers.py:149) and its two gradients, written so the A operand
// is read straight from the K separate basis tensors T_k[B*M, Fin] (no tf.stack /
// transpose round trip, gnn_layers.py:144-147) with the kernel's row order f*K + k.
// The same three kernels serve the pseudo-convolutions (Conv1D / Conv2DTranspose with
// kernel = stride = 4^p are plain GEMMs over NESTED-contiguous children).
//
// These are the parity-mode (rel <= 1e-5) kernels; the tensor-core modes live in
// ds_umma.cu.  64x64x16 tiles, 256 threads, 4x4 register tile per thread.
#include "ds_common.cuh"

namespace ds {
namespace {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

__device__ __forceinline__ void tile_fma(const float (*As)[BM + PAD], const float (*Bs)[BN + PAD], int ty, int tx,
                                         float acc[4][4]) {
#pragma unroll
  for (int kk = 0; kk < BK; ++kk) {
    const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
    const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
    const float av[4] = {a.x, a.y, a.z, a.w};
    const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}

// C[R,N] = act( sum_seg A_seg[R,Kc] * Bm[kc*bks + seg*bss, :] + bias )
__global__ void __launch_bounds__(256) gemm_nn_kernel(int64_t R, int64_t N, int64_t Kc, int nseg,
                                                      const float* __restrict__ A0, const float* __restrict__ Arest,
                                                      int64_t a_seg_stride, int64_t lda, const float* __restrict__ Bm,
                                                      int64_t ldb, int64_t bks, int64_t bss,
                                                      const float* __restrict__ bias, int64_t bias_mod, int act,
                                                      float* __restrict__ C, int64_t ldc) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int64_t r0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
  float acc[4][4] = {};
  for (int seg = 0; seg < nseg; ++seg) {
    const float* A = seg == 0 ? A0 : Arest + (int64_t)(seg - 1) * a_seg_stride;
    for (int64_t k0 = 0; k0 < Kc; k0 += BK) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = tid + i * 256;
        const int row = e >> 4, kk = e & 15;
        const int64_t r = r0 + row, k = k0 + kk;
        As[kk][row] = (r < R && k < Kc) ? __ldg(A + r * lda + k) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = tid + i * 256;
        const int kk = e >> 6, n = e & 63;
        const int64_t k = k0 + kk, c = n0 + n;
        Bs[kk][n] = (k < Kc && c < N) ? __ldg(Bm + (k * bks + seg * bss) * ldb + c) : 0.f;
      }
      __syncthreads();
      tile_fma(As, Bs, ty, tx, acc);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = r0 + ty * 4 + i;
    if (r >= R) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t c = n0 + tx * 4 + j;
      if (c >= N) continue;
      float v = acc[i][j];
      if (bias != nullptr) v += __ldg(bias + (c % bias_mod));
      C[r * ldc + c] = act_apply(v, act);
    }
  }
}

// C_seg[R,Nc] = act( A[R,Kd] * Bm[nc*bns + seg*bss, :Kd]^T + bias ),  seg = blockIdx.z
__global__ void __launch_bounds__(256) gemm_nt_kernel(int64_t R, int64_t Nc, int64_t Kd, const float* __restrict__ A,
                                                      int64_t lda, const float* __restrict__ Bm, int64_t ldb,
                                                      int64_t bns, int64_t bss, const float* __restrict__ bias,
                                                      int64_t bias_mod, int act, float* __restrict__ C, int64_t ldc,
                                                      int64_t c_seg_stride) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int seg = blockIdx.z;
  const int64_t r0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
  float acc[4][4] = {};
  for (int64_t k0 = 0; k0 < Kd; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      const int row = e >> 4, kk = e & 15;
      const int64_t r = r0 + row, k = k0 + kk;
      As[kk][row] = (r < R && k < Kd) ? __ldg(A + r * lda + k) : 0.f;
      const int64_t c = n0 + row;
      Bs[kk][row] = (c < Nc && k < Kd) ? __ldg(Bm + (c * bns + seg * bss) * ldb + k) : 0.f;
    }
    __syncthreads();
    tile_fma(As, Bs, ty, tx, acc);
    __syncthreads();
  }
  float* Cs = C + (int64_t)seg * c_seg_stride;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = r0 + ty * 4 + i;
    if (r >= R) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t c = n0 + tx * 4 + j;
      if (c >= Nc) continue;
      float v = acc[i][j];
      if (bias != nullptr) v += __ldg(bias + (c % bias_mod));
      Cs[r * ldc + c] = act_apply(v, act);
    }
  }
}

// C[R,Nc] = sum_seg A_seg[R,Kd] * Bm[nc*bns + seg*bss, :Kd]^T   (segmented A, one output)
__global__ void __launch_bounds__(256) gemm_nt_seg_kernel(int64_t R, int64_t Nc, int64_t Kd, int nseg,
                                                          const float* __restrict__ A0, const float* __restrict__ Arest,
                                                          int64_t a_seg_stride, int64_t lda,
                                                          const float* __restrict__ Bm, int64_t ldb, int64_t bns,
                                                          int64_t bss, float* __restrict__ C, int64_t ldc) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int64_t r0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
  float acc[4][4] = {};
  for (int seg = 0; seg < nseg; ++seg) {
    const float* A = seg == 0 ? A0 : Arest + (int64_t)(seg - 1) * a_seg_stride;
    for (int64_t k0 = 0; k0 < Kd; k0 += BK) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = tid + i * 256;
        const int row = e >> 4, kk = e & 15;
        const int64_t r = r0 + row, k = k0 + kk;
        As[kk][row] = (r < R && k < Kd) ? __ldg(A + r * lda + k) : 0.f;
        const int64_t c = n0 + row;
        Bs[kk][row] = (c < Nc && k < Kd) ? __ldg(Bm + (c * bns + seg * bss) * ldb + k) : 0.f;
      }
      __syncthreads();
      tile_fma(As, Bs, ty, tx, acc);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = r0 + ty * 4 + i;
    if (r >= R) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t c = n0 + tx * 4 + j;
      if (c < Nc) C[r * ldc + c] = acc[i][j];
    }
  }
}

// partial[split][seg][kc, n] = sum_{r in split} A_seg[r,kc] * D[r,n]
__global__ void __launch_bounds__(256) gemm_tn_kernel(int64_t R, int64_t N, int64_t Kc, int nseg, int kc_tiles,
                                                      const float* __restrict__ A0, const float* __restrict__ Arest,
                                                      int64_t a_seg_stride, int64_t lda, const float* __restrict__ D,
                                                      int64_t ldd, float* __restrict__ partial, int64_t rows_per_split) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Ds[BK][BN + PAD];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int seg = blockIdx.x / kc_tiles;
  const int64_t kc0 = (int64_t)(blockIdx.x % kc_tiles) * BM, n0 = (int64_t)blockIdx.y * BN;
  const int64_t split = blockIdx.z;
  const int64_t rb = split * rows_per_split;
  const int64_t re = min(R, rb + rows_per_split);
  const float* A = seg == 0 ? A0 : Arest + (int64_t)(seg - 1) * a_seg_stride;
  float acc[4][4] = {};
  for (int64_t rr0 = rb; rr0 < re; rr0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      const int rr = e >> 6, c = e & 63;
      const int64_t r = rr0 + rr;
      As[rr][c] = (r < re && kc0 + c < Kc) ? __ldg(A + r * lda + kc0 + c) : 0.f;
      Ds[rr][c] = (r < re && n0 + c < N) ? __ldg(D + r * ldd + n0 + c) : 0.f;
    }
    __syncthreads();
    tile_fma(As, Ds, ty, tx, acc);
    __syncthreads();
  }
  // partial layout: [split][seg][Kc][N]
  float* P = partial + ((split * nseg + seg) * Kc) * N;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t kc = kc0 + ty * 4 + i;
    if (kc >= Kc) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t c = n0 + tx * 4 + j;
      if (c < N) P[kc * N + c] = acc[i][j];
    }
  }
}

__global__ void gemm_tn_reduce_kernel(int64_t N, int64_t Kc, int nseg, int64_t splits, const float* __restrict__ partial,
                                      float* __restrict__ C, int64_t ldc, int64_t cks, int64_t css) {
  const int64_t total = (int64_t)nseg * Kc * N;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int64_t sp = 0; sp < splits; ++sp) s += partial[sp * total + e];
    const int64_t n = e % N;
    const int64_t kc = (e / N) % Kc;
    const int64_t seg = e / (N * Kc);
    C[(kc * cks + seg * css) * ldc + n] = s;
  }
}

__global__ void act_backward_kernel(int64_t n, const float* __restrict__ y, const float* __restrict__ dy, int act,
                                    float* __restrict__ dz) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dz[i] = dy[i] * act_grad_from_y(__ldg(y + i), act);
}

constexpr int COLSUM_BLOCKS = 512;

// partial[blk][c] = sum of Z[r, c] over the block's rows (deterministic order)
__global__ void __launch_bounds__(256) colsum_partial_kernel(int64_t R, int64_t Ncols, const float* __restrict__ Z,
                                                             float* __restrict__ partial) {
  __shared__ float red[256];
  const int t = threadIdx.x;
  const int cpb = (int)min((int64_t)256, Ncols);
  const int rpi = 256 / cpb;
  const int cl = t % cpb, rr = t / cpb;
  const int64_t rows_per_block = (R + gridDim.x - 1) / gridDim.x;
  const int64_t rb = (int64_t)blockIdx.x * rows_per_block;
  const int64_t re = min(R, rb + rows_per_block);
  for (int64_t c0 = 0; c0 < Ncols; c0 += cpb) {
    const int64_t c = c0 + cl;
    float acc = 0.f;
    if (rr < rpi && c < Ncols)
      for (int64_t r = rb + rr; r < re; r += rpi) acc += __ldg(Z + r * Ncols + c);
    red[t] = acc;
    __syncthreads();
    if (rr == 0 && c < Ncols) {
      float s = 0.f;
      for (int q = 0; q < rpi; ++q) s += red[q * cpb + cl];
      partial[(int64_t)blockIdx.x * Ncols + c] = s;
    }
    __syncthreads();
  }
}

// out[f] = sum over blocks and folded columns (c = f, f + F, ...) of partial[b][c].  One WARP per output: the lanes split
// the nblk * (Ncols / F) terms (independent loads, fixed assignment and a fixed shuffle tree: deterministic).  (Round 1
// used ceil(F / 32) CTAs whose threads each walked 64 block partials serially: one or two SMs busy for ~0.3 ms per call,
// 15 % of a HealpyGCNN training step in the ncu launch list.)
__global__ void __launch_bounds__(256) colsum_final_kernel(int64_t Ncols, int64_t F, int nblk,
                                                            const float* __restrict__ partial, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t f = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (f >= F) return;
  const int64_t fold = (Ncols - f + F - 1) / F;  // columns f, f + F, ... < Ncols
  const int64_t terms = (int64_t)nblk * fold;
  float s = 0.f;
  for (int64_t i = lane; i < terms; i += 32) {
    const int64_t b = i / fold, j = i - b * fold;
    s += __ldg(partial + b * Ncols + f + j * F);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[f] = s;
}

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

inline int64_t tn_splits(int64_t R, int64_t Kc, int nseg, int64_t N) {
  const int64_t tiles = cdiv(Kc, BM) * nseg * cdiv(N, BN);
  int64_t splits = cdiv((int64_t)4 * num_sms(), tiles);
  splits = std::min<int64_t>(splits, cdiv(R, 256));  // at least 256 rows per split
  splits = std::max<int64_t>(1, std::min<int64_t>(splits, 1024));
  return splits;
}

}  // namespace

int launch_gemm_nn(int64_t R, int64_t N, int64_t Kc, int nseg, const float* A0, const float* Arest,
                   int64_t a_seg_stride, int64_t lda, const float* Bm, int64_t ldb, int64_t b_kc_stride,
                   int64_t b_seg_stride, const float* bias, int64_t bias_mod, int act, float* C, int64_t ldc,
                   cudaStream_t st) {
  if (narrow_rows_usable(N, Kc, nseg))  // few output columns, short reduction: stream it (ds_narrow.cu)
    return launch_narrow_rows(R, N, Kc, nseg, A0, Arest, a_seg_stride, lda, Bm, b_kc_stride * ldb, b_seg_stride * ldb, 1, bias,
                              bias_mod, act, C, ldc, st);
  dim3 grid((unsigned)cdiv(R, BM), (unsigned)cdiv(N, BN));
  DS_CHECK(cdiv(N, BN) < 65536, "gemm_nn: N too large");
  gemm_nn_kernel<<<grid, 256, 0, st>>>(R, N, Kc, nseg, A0, Arest, a_seg_stride, lda, Bm, ldb, b_kc_stride,
                                       b_seg_stride, bias, bias_mod > 0 ? bias_mod : 1, act, C, ldc);
  DS_LAUNCHED();
  return 0;
}

int launch_gemm_nt(int64_t R, int64_t Nc, int64_t Kd, int nseg, const float* A, int64_t lda, const float* Bm,
                   int64_t ldb, int64_t b_nc_stride, int64_t b_seg_stride, const float* bias, int64_t bias_mod,
                   int act, float* C, int64_t ldc, int64_t c_seg_stride, cudaStream_t st) {
  dim3 grid((unsigned)cdiv(R, BM), (unsigned)cdiv(Nc, BN), (unsigned)nseg);
  DS_CHECK(cdiv(Nc, BN) < 65536 && nseg < 65536, "gemm_nt: N or nseg too large");
  gemm_nt_kernel<<<grid, 256, 0, st>>>(R, Nc, Kd, A, lda, Bm, ldb, b_nc_stride, b_seg_stride, bias,
                                       bias_mod > 0 ? bias_mod : 1, act, C, ldc, c_seg_stride);
  DS_LAUNCHED();
  return 0;
}

int launch_gemm_nt_seg(int64_t R, int64_t Nc, int64_t Kd, int nseg, const float* A0, const float* Arest,
                       int64_t a_seg_stride, int64_t lda, const float* Bm, int64_t ldb, int64_t b_nc_stride,
                       int64_t b_seg_stride, float* C, int64_t ldc, cudaStream_t st) {
  if (narrow_rows_usable(Nc, Kd, nseg))  // B(seg, kd, c) = Bm[(c * b_nc_stride + seg * b_seg_stride) * ldb + kd]
    return launch_narrow_rows(R, Nc, Kd, nseg, A0, Arest, a_seg_stride, lda, Bm, 1, b_seg_stride * ldb, b_nc_stride * ldb,
                              nullptr, 1, DS_ACT_LINEAR, C, ldc, st);
  dim3 grid((unsigned)cdiv(R, BM), (unsigned)cdiv(Nc, BN));
  DS_CHECK(cdiv(Nc, BN) < 65536, "gemm_nt_seg: N too large");
  gemm_nt_seg_kernel<<<grid, 256, 0, st>>>(R, Nc, Kd, nseg, A0, Arest, a_seg_stride, lda, Bm, ldb, b_nc_stride,
                                           b_seg_stride, C, ldc);
  DS_LAUNCHED();
  return 0;
}

int64_t gemm_tn_workspace_elems(int64_t R, int64_t Kc, int nseg, int64_t N) {
  return std::max(tn_splits(R, Kc, nseg, N) * nseg * Kc * N, narrow_tn_workspace_elems(R, Kc, nseg, N));
}

int64_t colsum_workspace_elems(int64_t Ncols) { return (int64_t)COLSUM_BLOCKS * Ncols; }

int launch_gemm_tn(int64_t R, int64_t N, int64_t Kc, int nseg, const float* A0, const float* Arest,
                   int64_t a_seg_stride, int64_t lda, const float* D, int64_t ldd, float* C, int64_t ldc,
                   int64_t c_kc_stride, int64_t c_seg_stride, float* partial, cudaStream_t st) {
  if (narrow_tn_usable(N, Kc, nseg)) {  // a [nseg * Kc, N] output of a few hundred elements: streaming kernel (ds_narrow.cu)
    int64_t nblk = 0;
    DS_TRY(launch_narrow_tn(R, N, Kc, nseg, A0, Arest, a_seg_stride, lda, D, ldd, partial, &nblk, st));
    const int64_t tot = (int64_t)nseg * Kc * N;
    gemm_tn_reduce_kernel<<<(unsigned)std::min<int64_t>(cdiv(tot, 256), 4096), 256, 0, st>>>(
        N, Kc, nseg, nblk, partial, C, ldc, c_kc_stride, c_seg_stride);
    DS_LAUNCHED();
    return 0;
  }
  const int64_t splits = tn_splits(R, Kc, nseg, N);
  const int64_t rows_per_split = cdiv(cdiv(R, splits), BK) * BK;
  const int kc_tiles = (int)cdiv(Kc, BM);
  dim3 grid((unsigned)(kc_tiles * nseg), (unsigned)cdiv(N, BN), (unsigned)splits);
  DS_CHECK(cdiv(N, BN) < 65536, "gemm_tn: N too large");
  gemm_tn_kernel<<<grid, 256, 0, st>>>(R, N, Kc, nseg, kc_tiles, A0, Arest, a_seg_stride, lda, D, ldd, partial,
                                       rows_per_split);
  DS_LAUNCHED();
  const int64_t total = (int64_t)nseg * Kc * N;
  gemm_tn_reduce_kernel<<<(unsigned)std::min<int64_t>(cdiv(total, 256), 4096), 256, 0, st>>>(
      N, Kc, nseg, splits, partial, C, ldc, c_kc_stride, c_seg_stride);
  DS_LAUNCHED();
  return 0;
}

int launch_act_backward(int64_t R, int64_t Ncols, int64_t F, const float* y, const float* dy, int act, float* dz,
                        cudaStream_t st) {
  (void)F;
  const int64_t n = R * Ncols;
  const int64_t blocks = std::min<int64_t>(cdiv(n, 256), (int64_t)num_sms() * 32);
  act_backward_kernel<<<(unsigned)blocks, 256, 0, st>>>(n, y, dy, act, dz);
  DS_LAUNCHED();
  return 0;
}

// out[f] = sum over `nblk` block partials [nblk, Ncols] (and folded columns c = f, f + F, ...)
int launch_colsum_final(int64_t Ncols, int64_t F, int nblk, const float* partial, float* out, cudaStream_t st) {
  colsum_final_kernel<<<(unsigned)cdiv(F, 8), 256, 0, st>>>(Ncols, F, nblk, partial, out);
  DS_LAUNCHED();
  return 0;
}

int launch_colsum(int64_t R, int64_t Ncols, int64_t F, const float* Z, float* out, float* workspace, cudaStream_t st) {
  const int nblk = (int)std::max<int64_t>(1, std::min<int64_t>(COLSUM_BLOCKS, cdiv(R, 64)));
  colsum_partial_kernel<<<nblk, 256, 0, st>>>(R, Ncols, Z, workspace);
  DS_LAUNCHED();
  colsum_final_kernel<<<(unsigned)cdiv(F, 8), 256, 0, st>>>(Ncols, F, nblk, workspace, out);
  DS_LAUNCHED();
  return 0;
}

}  // namespace ds
