// C-ABI glue of the fused lattice recursion: attaching the tile plan to a ds_plan, and the forward basis /
// backward Clenshaw drivers that combine the lattice kernel (every tile) with the generic kernels on a compact
// sub-problem: the own pixels within H hops' reach of the 8 valence-3 vertices of the HEALPix tessellation, where the
// neighbourhood is not a lattice (15 pixels in each of the 24 tiles around them for H = 4; deepsphere/lattice.py works
// them out per pixel).  The lattice kernel writes those rows too - wrongly -, the scatter of the sub-problem, which
// follows it on the stream, overwrites them.
#include <vector>

#include "ds_lattice.cuh"

namespace ds {

struct LatticeAttachment {
  LatticeDev dev;
  int H = 0;
  // irregular rows: generic kernels on the closure (all rows within H hops of them)
  ds_plan* sub_plan = nullptr;   // L~ restricted to the closure (owned)
  int64_t n_closure = 0, n_own = 0;
  int32_t* closure_rows = nullptr;  // [n_closure] global row of each closure row
  int32_t* own_sub = nullptr;       // [n_own] closure rows whose results are exact and wanted (the irregular rows)
  PatchDev patch;  // optional: the same rows as connected patches, for the one-launch kernel of ds_patch.cu
  int64_t device_bytes = 0;
};

namespace {

// dst[b, j, :] = src[b, rows[j], :]        (float4 granularity)
__global__ void gather_rows_kernel(int64_t B, int64_t M, int64_t n, int FV, const int32_t* __restrict__ rows,
                                   const float4* __restrict__ src, float4* __restrict__ dst) {
  const int64_t total = B * n * FV;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % FV);
    const int64_t j = (e / FV) % n, b = e / (FV * n);
    dst[e] = __ldg(src + (b * M + rows[j]) * FV + c);
  }
}
// dst[t][b, closure_rows[own[j]], :] = src[t][b, own[j], :] for nt tensors laid out back to back (B counts them all:
// the tensors are [B/nt, rows, F] each and contiguous, so tensor t, sample b is plain sample t * (B/nt) + b)
__global__ void scatter_rows_kernel(int64_t B, int64_t M, int64_t n_closure, int64_t n_own, int FV,
                                    const int32_t* __restrict__ closure_rows, const int32_t* __restrict__ own,
                                    const float4* __restrict__ src, float4* __restrict__ dst) {
  const int64_t total = B * n_own * FV;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % FV);
    const int64_t j = (e / FV) % n_own, b = e / (FV * n_own);
    const int32_t s = own[j];
    dst[(b * M + closure_rows[s]) * FV + c] = src[(b * n_closure + s) * FV + c];
  }
}

inline unsigned grid_for(int64_t n) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 16));
}

}  // namespace

void lattice_free(LatticeAttachment* L) {
  if (!L) return;
  cudaFree(L->dev.pix);
  cudaFree(L->dev.w);
  cudaFree(L->closure_rows);
  cudaFree(L->own_sub);
  cudaFree(L->patch.row_ptr);
  cudaFree(L->patch.rows);
  cudaFree(L->patch.ell_col);
  cudaFree(L->patch.ell_val);
  cudaFree(L->patch.own_ptr);
  cudaFree(L->patch.own_local);
  if (L->sub_plan) ds_plan_destroy(L->sub_plan);
  delete L;
}

// can the fused path serve this call?
bool lattice_usable(const ds_plan* plan, int32_t K, int64_t B, int64_t F) {
  if (!plan->lattice || !plan->symmetric) return false;
  const LatticeAttachment* L = plan->lattice;
  if (L->H < 1 || K < 2 || F % 4 != 0) return false;  // K - 1 > H: chained passes (compute_basis, ds_api.cu)
  LatticeArgs a;
  int threads = 0, smem = 0;
  return lattice_configure(L->dev, B, plan->M, (int)F, a, &threads, &smem) == 0;
}

int lattice_halo(const ds_plan* plan) { return plan->lattice ? plan->lattice->H : 0; }

// generic (unfused) recursion on the closure sub-problem, used for the irregular rows
// steps: out_s = alpha_s * S cur + beta_s * old + gamma_s * add_s on gathered tensors; results scattered
static int sub_problem_recursion(const ds_plan* plan, int64_t B, int F, int nsteps, const float* in0,
                                 const float* const* add, float* const* out, const float* alpha, const float* beta,
                                 const float* gamma, cudaStream_t st) {
  const LatticeAttachment* L = plan->lattice;
  if (L->n_own == 0) return 0;
  const int FV = F / 4;
  const int64_t n = L->n_closure, A = B * n * F;
  float* ws = nullptr;  // cur, old, new, add
  DS_CUDA(cudaMallocAsync((void**)&ws, sizeof(float) * 4 * A, st));
  float *cur = ws, *old = ws + A, *nxt = ws + 2 * A, *addb = ws + 3 * A;
  auto gather = [&](const float* src, float* dst) {
    gather_rows_kernel<<<grid_for(B * n * FV), 256, 0, st>>>(B, plan->M, n, FV, L->closure_rows,
                                                            reinterpret_cast<const float4*>(src),
                                                            reinterpret_cast<float4*>(dst));
    g_launches.fetch_add(1);
  };
  gather(in0, cur);
  int rc = 0;
  for (int s = 1; s <= nsteps && rc == 0; ++s) {
    const float* addp = nullptr;
    if (add[s - 1] != nullptr) {
      gather(add[s - 1], addb);
      addp = addb;
    }
    rc = launch_spmm(L->sub_plan->fwd, B, F, cur, alpha[s - 1], beta[s - 1] != 0.f ? old : nullptr, beta[s - 1], addp,
                     gamma[s - 1], nxt, st);
    if (rc == 0 && out[s - 1] != nullptr) {
      scatter_rows_kernel<<<grid_for(B * L->n_own * FV), 256, 0, st>>>(
          B, plan->M, n, L->n_own, FV, L->closure_rows, L->own_sub, reinterpret_cast<const float4*>(nxt),
          reinterpret_cast<float4*>(out[s - 1]));
      g_launches.fetch_add(1);
    }
    float* t = old; old = cur; cur = nxt; nxt = t;
  }
  cudaFreeAsync(ws, st);
  if (rc == 0) DS_CUDA(cudaGetLastError());
  return rc;
}

// the whole recursion (lattice kernel + irregular sub-problem)
int lattice_recursion(const ds_plan* plan, int64_t B, int F, int nsteps, const float* in0, const float* const* add,
                      float* const* out, const float* alpha, const float* beta, const float* gamma, cudaStream_t st) {
  const LatticeAttachment* L = plan->lattice;
  DS_CHECK(L != nullptr && nsteps >= 1 && nsteps <= L->H && nsteps <= LAT_MAX_STEPS, "lattice_recursion: plan mismatch");
  LatticeArgs a;
  int threads = 0, smem = 0;
  DS_CHECK(lattice_configure(L->dev, B, plan->M, F, a, &threads, &smem) == 0, "lattice_recursion: cannot configure");
  a.nsteps = nsteps;
  a.in0 = in0;
  for (int s = 0; s < LAT_MAX_STEPS; ++s) {
    a.add[s] = s < nsteps ? add[s] : nullptr;
    a.out[s] = s < nsteps ? out[s] : nullptr;
    a.alpha[s] = s < nsteps ? alpha[s] : 0.f;
    a.beta[s] = s < nsteps ? beta[s] : 0.f;
    a.gamma[s] = s < nsteps ? gamma[s] : 0.f;
  }
  DS_TRY(launch_lattice(L->dev, a, threads, smem, st));
  return sub_problem_recursion(plan, B, F, nsteps, in0, add, out, alpha, beta, gamma, st);
}

int umma_supported(int64_t Kc, int nseg, int64_t N);
int launch_umma_gemm(int64_t R, int64_t N, int64_t Kc, int nseg, const float* A0, const float* Arest,
                     int64_t a_seg_stride_rows, const float* Bm, int64_t b_k_stride, int64_t b_seg_stride,
                     int64_t b_n_stride, const float* bias, int64_t bias_mod, int act, float* C, int64_t ldc, int mode,
                     cudaStream_t st);

bool lattice_conv2_usable(const LatticeDev& L, int nsteps, int F, int N, int mode);
int launch_lattice_conv2(const LatticeDev& L, int nsteps, int64_t B, int64_t M, int F, int N, int recursion,
                         const float* in0, float* const* out, const float* W, int64_t s_f, int64_t s_k, int64_t s_n,
                         const float* bias, int act, float* y, cudaStream_t st);

bool fused_conv_usable(const ds_plan* plan, int32_t K, int64_t B, int64_t F, int64_t N, int32_t mode) {
  (void)B;
  if (!plan->lattice || !plan->symmetric || K < 2) return false;
  const LatticeAttachment* L = plan->lattice;
  if (umma_supported(F, K, N) != 0) return false;  // the irregular rows go through the tensor-core GEMM
  return lattice_conv2_usable(L->dev, K - 1, (int)F, (int)N, mode);  // register-resident kernel (ds_lattice_conv2.cu)
}

// y = act( sum_k T_k(L~)(in0) B_k + bias ),  B_k(f, n) = W[f*s_f + k*s_k + n*s_n];  basis_out (optional):
// [K-1, B, M, F] receives T_1..T_{K-1}(in0).  All tiles: one fused kernel; then the irregular rows: generic hops +
// tensor-core GEMM on the gathered closure, scattered over what the fused kernel left there.
int fused_conv(const ds_plan* plan, int32_t recursion, int32_t K, int64_t B, int64_t F, int64_t N, const float* in0,
               float* basis_out, const float* W, int64_t s_f, int64_t s_k, int64_t s_n, const float* bias, int act,
               float* y, int mode, cudaStream_t st) {
  const LatticeAttachment* L = plan->lattice;
  DS_CHECK(L != nullptr, "fused_conv: no lattice attachment");
  const int64_t M = plan->M, A = B * M * F;
  float* out[LAT_MAX_STEPS] = {};
  if (basis_out != nullptr)
    for (int s = 1; s < K; ++s) out[s - 1] = basis_out + (int64_t)(s - 1) * A;
  DS_CHECK(lattice_conv2_usable(L->dev, K - 1, (int)F, (int)N, mode), "fused_conv: shape / mode not served by the fused kernel");
  DS_TRY(launch_lattice_conv2(L->dev, K - 1, B, M, (int)F, (int)N, recursion, in0, basis_out ? out : nullptr, W, s_f,
                              s_k, s_n, bias, act, y, st));
  if (L->n_own == 0) return 0;
  // ---- irregular rows ----
  if (patch_usable(L->patch, K - 1, (int)F, (int)N))  // one launch: gather, hops, fp32 contraction, epilogue, stores
    return launch_patch_conv(L->patch, K - 1, B, M, (int)F, (int)N, recursion, in0, basis_out ? out : nullptr, W, s_f, s_k,
                             s_n, bias, act, y, st);
  const int FV = (int)(F / 4), NV = (int)(N / 4);
  const int64_t n = L->n_closure, As = B * n * F;
  float* ws = nullptr;  // [K][B, n, F] basis on the closure, then [B, n, N] result
  DS_CUDA(cudaMallocAsync((void**)&ws, sizeof(float) * ((int64_t)K * As + B * n * N), st));
  float* ys = ws + (int64_t)K * As;
  gather_rows_kernel<<<grid_for(B * n * FV), 256, 0, st>>>(B, M, n, FV, L->closure_rows,
                                                          reinterpret_cast<const float4*>(in0),
                                                          reinterpret_cast<float4*>(ws));
  g_launches.fetch_add(1);
  int rc = 0;
  for (int k = 1; k < K && rc == 0; ++k) {
    const bool cheb2 = recursion == DS_RECURSION_CHEBYSHEV && k >= 2;
    rc = launch_spmm(L->sub_plan->fwd, B, F, ws + (int64_t)(k - 1) * As, cheb2 ? 2.f : 1.f,
                     cheb2 ? ws + (int64_t)(k - 2) * As : nullptr, -1.f, nullptr, 0.f, ws + (int64_t)k * As, st);
  }
  if (rc == 0 && basis_out != nullptr) {  // T_1 .. T_{K-1}: [K-1][B, n, F] -> [K-1][B, M, F], one launch
    const int64_t BK = (int64_t)(K - 1) * B;
    scatter_rows_kernel<<<grid_for(BK * L->n_own * FV), 256, 0, st>>>(
        BK, M, n, L->n_own, FV, L->closure_rows, L->own_sub, reinterpret_cast<const float4*>(ws + As),
        reinterpret_cast<float4*>(basis_out));
    g_launches.fetch_add(1);
  }
  if (rc == 0)
    rc = launch_umma_gemm(B * n, N, F, K, ws, ws + As, B * n, W, s_f, s_k, s_n, bias, N, act, ys, N, mode, st);
  if (rc == 0) {
    scatter_rows_kernel<<<grid_for(B * L->n_own * NV), 256, 0, st>>>(B, M, n, L->n_own, NV, L->closure_rows, L->own_sub,
                                                                     reinterpret_cast<const float4*>(ys),
                                                                     reinterpret_cast<float4*>(y));
    g_launches.fetch_add(1);
  }
  cudaFreeAsync(ws, st);
  if (rc == 0) DS_CUDA(cudaGetLastError());
  return rc;
}

}  // namespace ds

extern "C" int ds_plan_attach_lattice(ds_plan_t* plan, int32_t n_tiles, int32_t LW, int32_t H, int32_t T,
                                      const int32_t* pix, const float* w, ds_plan_t* sub_plan, int64_t n_closure,
                                      const int32_t* closure_rows, int64_t n_own, const int32_t* own_sub) {
  using namespace ds;
  DS_CHECK(plan != nullptr, "ds_plan_attach_lattice: NULL plan");
  DS_CHECK(n_tiles >= 0 && LW == T + 2 * H && H >= 1 && H <= LAT_MAX_STEPS && T >= 4, "ds_plan_attach_lattice: bad geometry");
  DS_CHECK(n_tiles == 0 || (pix && w), "ds_plan_attach_lattice: NULL tables");
  DS_CHECK(n_own == 0 || (sub_plan && closure_rows && own_sub && sub_plan->M == n_closure),
           "ds_plan_attach_lattice: inconsistent irregular sub-problem");
  if (plan->lattice) {
    lattice_free(plan->lattice);
    plan->lattice = nullptr;
  }
  LatticeAttachment* L = new LatticeAttachment();
  L->H = H;
  L->dev.n_tiles = n_tiles; L->dev.LW = LW; L->dev.H = H; L->dev.T = T;
  const size_t P = (size_t)LW * LW;
  {  // is the diagonal of L~ one scalar on every position that carries a pixel?  (holes: all nine weights are zero)
    bool first = true, same = true;
    float d0 = 0.f;
    for (size_t i = 0; i < (size_t)n_tiles * P && same; ++i) {
      if (pix[i] < 0) continue;
      const float d = w[i * 9 + 8];
      if (first) { d0 = d; first = false; }
      else if (d != d0) same = false;
    }
    L->dev.diag_const = same && !first;
    L->dev.diag = d0;
  }
  auto up = [&](const void* src, size_t bytes, void** dst) -> int {
    *dst = nullptr;
    DS_CUDA(cudaMalloc(dst, std::max<size_t>(bytes, 16)));
    if (bytes) DS_CUDA(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
    L->device_bytes += (int64_t)bytes;
    return 0;
  };
  int rc = up(pix, sizeof(int32_t) * n_tiles * P, (void**)&L->dev.pix);
  if (rc == 0) rc = up(w, sizeof(float) * n_tiles * P * 9, (void**)&L->dev.w);
  L->n_closure = n_closure;
  L->n_own = n_own;
  if (rc == 0) rc = up(closure_rows, sizeof(int32_t) * n_closure, (void**)&L->closure_rows);
  if (rc == 0) rc = up(own_sub, sizeof(int32_t) * n_own, (void**)&L->own_sub);
  if (rc != 0) {
    lattice_free(L);
    return rc;
  }
  L->sub_plan = n_own > 0 ? sub_plan : nullptr;  // ownership moves to the parent plan
  plan->lattice = L;
  plan->device_bytes += L->device_bytes;
  return 0;
}

extern "C" int ds_plan_attach_patches(ds_plan_t* plan, int32_t n_patches, const int32_t* row_ptr, const int32_t* rows,
                                      const int32_t* ell_col, const float* ell_val, const int32_t* own_ptr,
                                      const int32_t* own_local) {
  using namespace ds;
  DS_CHECK(plan != nullptr && plan->lattice != nullptr, "ds_plan_attach_patches: the plan has no lattice attachment");
  DS_CHECK(n_patches >= 1 && row_ptr && rows && ell_col && ell_val && own_ptr && own_local,
           "ds_plan_attach_patches: bad argument");
  LatticeAttachment* L = plan->lattice;
  DS_CHECK(L->patch.n_patches == 0, "ds_plan_attach_patches: already attached");
  const int64_t n_rows = row_ptr[n_patches], n_own = own_ptr[n_patches];
  DS_CHECK(row_ptr[0] == 0 && own_ptr[0] == 0 && n_rows == L->n_closure && n_own == L->n_own,
           "ds_plan_attach_patches: the patches must partition the closure and list every irregular row");
  PatchDev P;
  for (int p = 0; p < n_patches; ++p) {
    const int nr = row_ptr[p + 1] - row_ptr[p], no = own_ptr[p + 1] - own_ptr[p];
    DS_CHECK(nr >= 1 && no >= 0, "ds_plan_attach_patches: empty patch");
    for (int i = own_ptr[p]; i < own_ptr[p + 1]; ++i)
      DS_CHECK(own_local[i] >= 0 && own_local[i] < nr, "ds_plan_attach_patches: wanted row outside its patch");
    for (int64_t e = (int64_t)row_ptr[p] * 9; e < (int64_t)row_ptr[p + 1] * 9; ++e)
      DS_CHECK(ell_col[e] >= -1 && ell_col[e] < nr, "ds_plan_attach_patches: column outside its patch");
    P.max_rows = std::max(P.max_rows, nr);
    P.max_own = std::max(P.max_own, no);
  }
  for (int64_t i = 0; i < n_rows; ++i)
    DS_CHECK(rows[i] >= 0 && rows[i] < plan->M, "ds_plan_attach_patches: row outside the graph");
  int64_t bytes = 0;
  auto up = [&](const void* src, size_t nbytes, void** dst) -> int {
    *dst = nullptr;
    DS_CUDA(cudaMalloc(dst, std::max<size_t>(nbytes, 16)));
    if (nbytes) DS_CUDA(cudaMemcpy(*dst, src, nbytes, cudaMemcpyHostToDevice));
    bytes += (int64_t)nbytes;
    return 0;
  };
  int rc = up(row_ptr, sizeof(int32_t) * (n_patches + 1), (void**)&P.row_ptr);
  if (rc == 0) rc = up(rows, sizeof(int32_t) * n_rows, (void**)&P.rows);
  if (rc == 0) rc = up(ell_col, sizeof(int32_t) * n_rows * 9, (void**)&P.ell_col);
  if (rc == 0) rc = up(ell_val, sizeof(float) * n_rows * 9, (void**)&P.ell_val);
  if (rc == 0) rc = up(own_ptr, sizeof(int32_t) * (n_patches + 1), (void**)&P.own_ptr);
  if (rc == 0) rc = up(own_local, sizeof(int32_t) * n_own, (void**)&P.own_local);
  if (rc != 0) {
    cudaFree(P.row_ptr); cudaFree(P.rows); cudaFree(P.ell_col); cudaFree(P.ell_val); cudaFree(P.own_ptr);
    cudaFree(P.own_local);
    return rc;
  }
  P.n_patches = n_patches;
  L->patch = P;
  L->device_bytes += bytes;
  plan->device_bytes += bytes;
  return 0;
}
