// Streaming kernels for NARROW feature contractions: the layers the reference actually ships have 1 - 5 channels
// (examples/quick_start.ipynb:118-127: K = 10, Fout = 5; advanced_tutorial.ipynb:309-325), so the K*Fin -> Fout
// contraction of gnn_layers.py:149 and its two gradients are [R, 10..50] x [10..50, 5] products with R = B*M rows in the
// millions.  The 64 x 64 x 16 tiles of ds_gemm.cu spend such a problem on padding (ncu launch list of the quick_start
// training step, round 2: gemm_tn_kernel 0.77 ms per call for a 50 x 5 output = 35 % of the step, gemm_nn_kernel 0.24 ms);
// these kernels stream it:
//   narrow_rows_kernel   one thread per row r: C[r, :] = act( sum_{seg, k} A_seg[r, k] * B(seg, k, :) + bias ) with the whole
//                        B operand in shared memory (warp-broadcast reads); serves the forward (NN) and, with other strides,
//                        the data gradient (NT, segmented A)
//   narrow_tn_kernel     weight gradient: one thread per output element (seg, k, n), a tile of rows staged in shared memory,
//                        block partials reduced in a fixed order (the existing gemm_tn_reduce_kernel): deterministic
// Exact fp32 FMA chains in ascending (seg, k) order, like the tiled kernels: same rounding.
#include <algorithm>
#include <cstdlib>

#include "ds_common.cuh"

namespace ds {
namespace {

constexpr int NR_THREADS = 256;
constexpr int NR_NMAX = 16;   // outputs per row held in registers
constexpr int TN_ROWS = 64;   // rows per shared-memory tile of the weight-gradient kernel
constexpr int TN_OPT = 4;     // outputs per thread (256 threads: <= 1024 outputs)

// B(seg, k, n) = Bm[k * sk + seg * ss + n * sn]
__global__ void __launch_bounds__(NR_THREADS) narrow_rows_kernel(int64_t R, int N, int K, int nseg,
                                                                 const float* __restrict__ A0, const float* __restrict__ Arest,
                                                                 int64_t a_seg_stride, int64_t lda,
                                                                 const float* __restrict__ Bm, int64_t sk, int64_t ss, int64_t sn,
                                                                 const float* __restrict__ bias, int bias_mod, int act,
                                                                 float* __restrict__ C, int64_t ldc) {
  extern __shared__ float nr_Bs[];  // [(seg * K + k) * N + n]
  const int nb = nseg * K * N;
  for (int e = threadIdx.x; e < nb; e += NR_THREADS) {
    const int n = e % N, k = (e / N) % K, seg = e / (N * K);
    nr_Bs[e] = __ldg(Bm + (int64_t)k * sk + (int64_t)seg * ss + (int64_t)n * sn);
  }
  __syncthreads();
  for (int64_t r = (int64_t)blockIdx.x * NR_THREADS + threadIdx.x; r < R; r += (int64_t)gridDim.x * NR_THREADS) {
    float acc[NR_NMAX];
#pragma unroll
    for (int n = 0; n < NR_NMAX; ++n) acc[n] = 0.f;
    const float* bs = nr_Bs;
    for (int seg = 0; seg < nseg; ++seg) {
      const float* A = (seg == 0 ? A0 : Arest + (int64_t)(seg - 1) * a_seg_stride) + r * lda;
      for (int k = 0; k < K; ++k) {
        const float a = __ldg(A + k);
#pragma unroll
        for (int n = 0; n < NR_NMAX; ++n)
          if (n < N) acc[n] = fmaf(a, bs[n], acc[n]);
        bs += N;
      }
    }
    float* c = C + r * ldc;
#pragma unroll
    for (int n = 0; n < NR_NMAX; ++n)
      if (n < N) {
        float v = acc[n];
        if (bias != nullptr) v += __ldg(bias + (n % bias_mod));
        c[n] = act_apply(v, act);
      }
  }
}

// partial[block][(seg * K + k) * N + n] = sum over the block's rows of A_seg[r, k] * D[r, n]
__global__ void __launch_bounds__(NR_THREADS) narrow_tn_kernel(int64_t R, int N, int K, int nseg,
                                                               const float* __restrict__ A0, const float* __restrict__ Arest,
                                                               int64_t a_seg_stride, int64_t lda, const float* __restrict__ D,
                                                               int64_t ldd, float* __restrict__ partial, int64_t rows_per_block) {
  extern __shared__ float tn_smem[];
  const int KA = nseg * K;                 // A values per row
  float* As = tn_smem;                     // [TN_ROWS][KA]
  float* Ds = tn_smem + TN_ROWS * KA;      // [TN_ROWS][N]
  const int O = KA * N;
  const int64_t rb = (int64_t)blockIdx.x * rows_per_block;
  const int64_t re = min(R, rb + rows_per_block);
  float acc[TN_OPT];
  int ja[TN_OPT], jn[TN_OPT];
#pragma unroll
  for (int i = 0; i < TN_OPT; ++i) {
    const int o = threadIdx.x + i * NR_THREADS;
    acc[i] = 0.f;
    ja[i] = o < O ? o / N : 0;   // seg * K + k
    jn[i] = o < O ? o % N : 0;
  }
  for (int64_t r0 = rb; r0 < re; r0 += TN_ROWS) {
    const int rows = (int)min((int64_t)TN_ROWS, re - r0);
    for (int e = threadIdx.x; e < rows * KA; e += NR_THREADS) {
      const int rr = e / KA, j = e - rr * KA, seg = j / K, k = j - seg * K;
      const float* A = seg == 0 ? A0 : Arest + (int64_t)(seg - 1) * a_seg_stride;
      As[e] = __ldg(A + (r0 + rr) * lda + k);
    }
    for (int e = threadIdx.x; e < rows * N; e += NR_THREADS) {
      const int rr = e / N, n = e - rr * N;
      Ds[e] = __ldg(D + (r0 + rr) * ldd + n);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < TN_OPT; ++i) {
      if (threadIdx.x + i * NR_THREADS < O) {
        float s = acc[i];
        for (int rr = 0; rr < rows; ++rr) s = fmaf(As[rr * KA + ja[i]], Ds[rr * N + jn[i]], s);
        acc[i] = s;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TN_OPT; ++i) {
    const int o = threadIdx.x + i * NR_THREADS;
    if (o < O) partial[(int64_t)blockIdx.x * O + o] = acc[i];
  }
}

inline bool narrow_enabled() {
  static const bool on = [] { const char* e = getenv("DEEPSPHERE_NARROW"); return e == nullptr || atoi(e) != 0; }();
  return on;
}

}  // namespace

// narrow = few output columns AND a short reduction per segment: the tiled kernels would run on > 90 % padding
bool narrow_rows_usable(int64_t N, int64_t K, int nseg) {
  return narrow_enabled() && N >= 1 && N <= NR_NMAX && K >= 1 && K <= 16 && nseg >= 1 && (int64_t)nseg * K * N * 4 <= 48 * 1024;
}
bool narrow_tn_usable(int64_t N, int64_t K, int nseg) {
  return narrow_enabled() && N >= 1 && N <= 16 && K >= 1 && K <= 16 && (int64_t)nseg * K * N <= NR_THREADS * TN_OPT &&
         (int64_t)TN_ROWS * (nseg * K + N) * 4 <= 48 * 1024;
}
int64_t narrow_tn_blocks(int64_t R) {
  return std::max<int64_t>(1, std::min<int64_t>((int64_t)num_sms() * 4, (R + 4 * TN_ROWS - 1) / (4 * TN_ROWS)));
}
int64_t narrow_tn_workspace_elems(int64_t R, int64_t K, int nseg, int64_t N) {
  return narrow_tn_usable(N, K, nseg) ? narrow_tn_blocks(R) * nseg * K * N : 0;
}

int launch_narrow_rows(int64_t R, int64_t N, int64_t K, int nseg, const float* A0, const float* Arest, int64_t a_seg_stride,
                       int64_t lda, const float* Bm, int64_t sk, int64_t ss, int64_t sn, const float* bias, int64_t bias_mod,
                       int act, float* C, int64_t ldc, cudaStream_t st) {
  const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((R + NR_THREADS - 1) / NR_THREADS, (int64_t)num_sms() * 8));
  narrow_rows_kernel<<<blocks, NR_THREADS, (size_t)nseg * K * N * 4, st>>>(
      R, (int)N, (int)K, nseg, A0, Arest, a_seg_stride, lda, Bm, sk, ss, sn, bias, (int)(bias_mod > 0 ? bias_mod : 1), act, C, ldc);
  DS_LAUNCHED();
  return 0;
}

// writes partial[blocks][nseg * K * N]; returns the number of blocks through *nblk (the caller reduces them)
int launch_narrow_tn(int64_t R, int64_t N, int64_t K, int nseg, const float* A0, const float* Arest, int64_t a_seg_stride,
                     int64_t lda, const float* D, int64_t ldd, float* partial, int64_t* nblk, cudaStream_t st) {
  const int64_t blocks = narrow_tn_blocks(R);
  const int64_t rows_per_block = ((R + blocks - 1) / blocks + TN_ROWS - 1) / TN_ROWS * TN_ROWS;
  const int64_t used = (R + rows_per_block - 1) / rows_per_block;
  narrow_tn_kernel<<<(unsigned)used, NR_THREADS, (size_t)TN_ROWS * (nseg * K + N) * 4, st>>>(
      R, (int)N, (int)K, nseg, A0, Arest, a_seg_stride, lda, D, ldd, partial, rows_per_block);
  DS_LAUNCHED();
  *nblk = used;
  return 0;
}

}  // namespace ds
