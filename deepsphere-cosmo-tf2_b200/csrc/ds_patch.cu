// patch_conv_kernel: the whole graph convolution y = act( sum_k T_k(L~) x  B_k + bias ) on the IRREGULAR ROWS of a lattice
// plan in one launch (Chebyshev.call / Monomial.call, gnn_layers.py:131-152 / :283-299 of the reference, restricted to
// those rows).
//
// The fused lattice kernel is exact everywhere except on the own pixels within H hops' reach of the 8 valence-3 vertices
// of the HEALPix tessellation (deepsphere/lattice.py: 15 pixels in each of the 24 tiles around them for H = 4).  Their H-hop
// neighbourhoods are 8 disconnected patches of 189 rows.  One CTA takes one (patch, batch element): the patch's rows of x
// are gathered into shared memory, the K - 1 hops run there on the patch's own 9-wide ELL (three rotating [rows][F]
// buffers), and after every hop the wanted rows are contracted with B_k straight out of shared memory (fp32 FMA, the
// accumulators stay in registers over the hops).  Before this kernel the same rows cost 8 (forward) / 9 (backward-data)
// launches per layer - gather, K - 1 generic hops, weight image, tensor-core GEMM, scatters - each far below the size
// at which a launch is more than its latency (profiles/r2p_launches_model_train.csv: 33 spmm_tile + 18 scatter launches
// per HealpyGCNN step).
//
// Bound: launch latency / one wave of <= n_patches * B CTAs; traffic = the patch rows of x in, the wanted rows of y
// (and T_1..T_{K-1} when the caller wants the basis) out.
#include "ds_lattice.cuh"

namespace ds {

constexpr int PATCH_THREADS = 256;
constexpr int PATCH_MAXACC = 16;    // wanted rows per thread (registers)
constexpr int PATCH_MAX_STEPS = 4;  // hops in one launch (basis output pointers)
constexpr int PATCH_ELL = 9;        // 8 neighbours + diagonal

struct PatchArgs {
  const int32_t *row_ptr, *rows, *ell_col, *own_ptr, *own_local;
  const float* ell_val;
  int64_t B, M;
  int F, N, nsteps, cheb, act, max_rows;
  const float* in0;             // [B, M, F]
  float* out[PATCH_MAX_STEPS];  // optional T_1..T_nsteps [B, M, F] (wanted rows only)
  const float* W;               // B_k(f, n) = W[f*s_f + k*s_k + n*s_n]
  int64_t s_f, s_k, s_n;
  const float* bias;            // [N] or NULL
  float* y;                     // [B, M, N] (wanted rows only)
};

__host__ __device__ inline size_t patch_smem_bytes(int max_rows, int F, int N) {
  return (size_t)3 * max_rows * F * 4 + (size_t)max_rows * PATCH_ELL * 8 + (size_t)F * (N + 1) * 4 + 16;
}

#ifndef DS_EMULATE
#define PATCH_DYNAMIC_SMEM(name) extern __shared__ __align__(16) uint8_t name[]
#else
#define PATCH_DYNAMIC_SMEM(name) uint8_t* const name = emul::dynamic_smem()
#endif

__global__ void __launch_bounds__(PATCH_THREADS) patch_conv_kernel(const PatchArgs a) {
  PATCH_DYNAMIC_SMEM(smem);
  const int tid = (int)threadIdx.x;
  const int p = (int)(blockIdx.x / a.B);
  const int64_t b = blockIdx.x % a.B;
  const int r0 = a.row_ptr[p], nr = a.row_ptr[p + 1] - r0;
  const int o0 = a.own_ptr[p], no = a.own_ptr[p + 1] - o0;
  const int F = a.F, N = a.N, FV = F / 4;

  float4* cur = reinterpret_cast<float4*>(smem);
  float4* old = cur + (size_t)a.max_rows * FV;
  float4* nxt = old + (size_t)a.max_rows * FV;
  int32_t* scol = reinterpret_cast<int32_t*>(nxt + (size_t)a.max_rows * FV);
  float* sval = reinterpret_cast<float*>(scol + (size_t)a.max_rows * PATCH_ELL);
  float* sW = sval + (size_t)a.max_rows * PATCH_ELL;  // [F][N + 1]: one B_k at a time
  const int ldw = N + 1;

  for (int i = tid; i < nr * PATCH_ELL; i += PATCH_THREADS) {
    scol[i] = a.ell_col[(size_t)r0 * PATCH_ELL + i];
    sval[i] = a.ell_val[(size_t)r0 * PATCH_ELL + i];
  }
  const float4* x4 = reinterpret_cast<const float4*>(a.in0);
  for (int e = tid; e < nr * FV; e += PATCH_THREADS) {
    const int r = e / FV, c = e % FV;
    cur[e] = __ldg(x4 + ((size_t)b * a.M + a.rows[r0 + r]) * FV + c);
  }
  __syncthreads();  // T_0 and the ELL

  // contraction roles: column n = tid % N for rows ty, ty + RY, ... of the wanted list (RY = 256 / N row groups)
  const int RY = PATCH_THREADS / N;
  const int n = tid % N, ty = tid / N;
  const bool active = ty < RY;
  int lrow[PATCH_MAXACC];  // patch-local row of each accumulator (-1: none)
  float acc[PATCH_MAXACC];
#pragma unroll
  for (int q = 0; q < PATCH_MAXACC; ++q) {
    const int i = ty + q * RY;
    lrow[q] = (active && i < no) ? a.own_local[o0 + i] : -1;
    acc[q] = 0.f;
  }

  for (int k = 0;; ++k) {
    // ---- stage B_k as [F][N + 1] (the padded pitch keeps the transposed fill of a backward launch conflict-free) ----
    if (a.s_n == 1 || a.s_f != 1) {
      for (int e = tid; e < F * N; e += PATCH_THREADS) {
        const int f = e / N, nn = e % N;
        sW[f * ldw + nn] = __ldg(a.W + (size_t)f * a.s_f + (size_t)k * a.s_k + (size_t)nn * a.s_n);
      }
    } else {
      for (int e = tid; e < F * N; e += PATCH_THREADS) {
        const int nn = e / F, f = e % F;
        sW[f * ldw + nn] = __ldg(a.W + (size_t)f * a.s_f + (size_t)k * a.s_k + (size_t)nn * a.s_n);
      }
    }
    // ---- hop k -> k + 1 on the patch ----
    if (k < a.nsteps) {
      const bool cheb2 = a.cheb != 0 && k >= 1;
      for (int e = tid; e < nr * FV; e += PATCH_THREADS) {
        const int r = e / FV, c = e % FV;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < PATCH_ELL; ++j) {
          const int col = scol[r * PATCH_ELL + j];
          if (col >= 0) {
            const float v = sval[r * PATCH_ELL + j];
            const float4 t = cur[col * FV + c];
            s.x = fmaf(v, t.x, s.x);
            s.y = fmaf(v, t.y, s.y);
            s.z = fmaf(v, t.z, s.z);
            s.w = fmaf(v, t.w, s.w);
          }
        }
        if (cheb2) {
          const float4 o = old[e];
          s.x = 2.f * s.x - o.x;
          s.y = 2.f * s.y - o.y;
          s.z = 2.f * s.z - o.z;
          s.w = 2.f * s.w - o.w;
        }
        nxt[e] = s;
      }
    }
    __syncthreads();
    // ---- acc += T_k[wanted rows] . B_k ----
    if (active) {
      for (int f4 = 0; f4 < FV; ++f4) {
        const float w0 = sW[(4 * f4 + 0) * ldw + n], w1 = sW[(4 * f4 + 1) * ldw + n];
        const float w2 = sW[(4 * f4 + 2) * ldw + n], w3 = sW[(4 * f4 + 3) * ldw + n];
#pragma unroll
        for (int q = 0; q < PATCH_MAXACC; ++q) {
          if (lrow[q] >= 0) {
            const float4 t = cur[lrow[q] * FV + f4];
            acc[q] = fmaf(t.x, w0, acc[q]);
            acc[q] = fmaf(t.y, w1, acc[q]);
            acc[q] = fmaf(t.z, w2, acc[q]);
            acc[q] = fmaf(t.w, w3, acc[q]);
          }
        }
      }
    }
    if (k == a.nsteps) break;
    // ---- T_{k+1} on the wanted rows for the caller (the weight gradient's operand) ----
    if (a.out[k] != nullptr) {
      float4* o4 = reinterpret_cast<float4*>(a.out[k]);
      for (int e = tid; e < no * FV; e += PATCH_THREADS) {
        const int i = e / FV, c = e % FV;
        const int lr = a.own_local[o0 + i];
        o4[((size_t)b * a.M + a.rows[r0 + lr]) * FV + c] = nxt[lr * FV + c];
      }
    }
    __syncthreads();  // every reader of cur / old / B_k is done before the next fill
    float4* t = old;
    old = cur;
    cur = nxt;
    nxt = t;
  }

  if (active) {
    const float bn = a.bias != nullptr ? __ldg(a.bias + n) : 0.f;
#pragma unroll
    for (int q = 0; q < PATCH_MAXACC; ++q)
      if (lrow[q] >= 0) a.y[((size_t)b * a.M + a.rows[r0 + lrow[q]]) * N + n] = act_apply(acc[q] + bn, a.act);
  }
}

#ifndef DS_EMULATE  // tests/emul runs the kernel above on the host; the launcher below needs nvcc

bool patch_usable(const PatchDev& P, int nsteps, int F, int N) {
  static const bool enabled = [] {
    const char* e = getenv("DEEPSPHERE_PATCH");
    return !(e != nullptr && e[0] == '0');
  }();
  if (!enabled || P.n_patches <= 0 || nsteps < 1 || nsteps > PATCH_MAX_STEPS) return false;
  if (F % 4 != 0 || N < 1 || N > PATCH_THREADS) return false;
  if ((int64_t)(PATCH_THREADS / N) * PATCH_MAXACC < P.max_own) return false;
  return patch_smem_bytes(P.max_rows, F, N) <= 227 * 1024;
}

int launch_patch_conv(const PatchDev& P, int nsteps, int64_t B, int64_t M, int F, int N, int recursion, const float* in0,
                      float* const* out, const float* W, int64_t s_f, int64_t s_k, int64_t s_n, const float* bias,
                      int act, float* y, cudaStream_t st) {
  DS_CHECK(patch_usable(P, nsteps, F, N), "patch_conv: shape not served");
  DS_CHECK((int64_t)P.n_patches * B < (int64_t)1 << 31, "patch_conv: grid too large");
  PatchArgs a;
  a.row_ptr = P.row_ptr; a.rows = P.rows; a.ell_col = P.ell_col; a.own_ptr = P.own_ptr; a.own_local = P.own_local;
  a.ell_val = P.ell_val;
  a.B = B; a.M = M; a.F = F; a.N = N; a.nsteps = nsteps; a.cheb = recursion == DS_RECURSION_CHEBYSHEV; a.act = act;
  a.max_rows = P.max_rows;
  a.in0 = in0;
  for (int s = 0; s < PATCH_MAX_STEPS; ++s) a.out[s] = (out != nullptr && s < nsteps) ? out[s] : nullptr;
  a.W = W; a.s_f = s_f; a.s_k = s_k; a.s_n = s_n; a.bias = bias; a.y = y;
  const size_t smem = patch_smem_bytes(P.max_rows, F, N);
  static PerDeviceOnce attr_once;
  DS_TRY(attr_once.run([&]() -> int {
    DS_CUDA(cudaFuncSetAttribute(patch_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    return 0;
  }));
  patch_conv_kernel<<<(unsigned)(P.n_patches * B), PATCH_THREADS, smem, st>>>(a);
  DS_LAUNCHED();
  DS_CUDA(cudaGetLastError());
  return 0;
}

#endif  // DS_EMULATE

}  // namespace ds
