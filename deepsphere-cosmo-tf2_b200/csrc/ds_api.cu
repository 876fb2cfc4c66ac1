// C-ABI compute entries: graph convolution forward/backward (Chebyshev / Monomial),
// bias+activation, pseudo-convolutions.  Each entry is a short sequence of launches of the
// kernels in ds_spmm.cu / ds_gemm.cu / ds_umma.cu on the caller's stream.
#include <algorithm>

#include "ds_common.cuh"

namespace ds {

// tensor-core contraction (ds_umma.cu).  Returns -1 when the shape is not supported by
// the tcgen05 path (caller then reports an error; there is no silent fallback).
int umma_supported(int64_t Kc, int nseg, int64_t N);
int launch_umma_gemm_nn(int64_t R, int64_t N, int64_t Kc, int nseg, const float* A0, const float* Arest,
                        int64_t a_seg_stride, const float* Bm, int64_t b_kc_stride, int64_t b_seg_stride,
                        const float* bias, int act, float* C, int mode, cudaStream_t st);
int launch_umma_gemm(int64_t R, int64_t N, int64_t Kc, int nseg, const float* A0, const float* Arest,
                     int64_t a_seg_stride_rows, const float* Bm, int64_t b_k_stride, int64_t b_seg_stride,
                     int64_t b_n_stride, const float* bias, int64_t bias_mod, int act, float* C, int64_t ldc, int mode,
                     cudaStream_t st);

int umma_tn_supported(int64_t R, int64_t N, int64_t Kc, int nseg);
int64_t umma_tn_workspace_elems(int64_t R, int64_t N, int64_t Kc, int nseg);
int launch_umma_gemm_tn(int64_t R, int64_t N, int64_t Kc, int nseg, const float* A0, const float* Arest, const float* D,
                        float* C, int64_t s_kc, int64_t s_seg, int64_t s_n, float* partial, int mode, cudaStream_t st);
// fully fused recursion + contraction on the lattice (ds_lattice_conv2.cu / ds_lattice_api.cu)
bool fused_conv_usable(const ds_plan* plan, int32_t K, int64_t B, int64_t F, int64_t N, int32_t mode);
int fused_conv(const ds_plan* plan, int32_t recursion, int32_t K, int64_t B, int64_t F, int64_t N, const float* in0,
               float* basis_out, const float* W, int64_t s_f, int64_t s_k, int64_t s_n, const float* bias, int act,
               float* y, int mode, cudaStream_t st);

bool lattice_usable(const ds_plan* plan, int32_t K, int64_t B, int64_t F);
int lattice_halo(const ds_plan* plan);
int lattice_recursion(const ds_plan* plan, int64_t B, int F, int nsteps, const float* in0, const float* const* add,
                      float* const* out, const float* alpha, const float* beta, const float* gamma, cudaStream_t st);

namespace {

int check_common(const ds_plan_t* plan, int32_t recursion, int32_t K, int64_t B, int64_t Fin, int64_t Fout,
                 int32_t act, int32_t mode, const char* who) {
  DS_CHECK(plan != nullptr, "%s: NULL plan", who);
  DS_CHECK(recursion == DS_RECURSION_CHEBYSHEV || recursion == DS_RECURSION_MONOMIAL, "%s: bad recursion %d", who,
           recursion);
  DS_CHECK(K >= 1, "%s: K must be >= 1 (got %d)", who, K);
  DS_CHECK(B >= 1 && Fin >= 1 && Fout >= 1, "%s: empty shape B=%lld Fin=%lld Fout=%lld", who, (long long)B,
           (long long)Fin, (long long)Fout);
  DS_CHECK(act >= DS_ACT_LINEAR && act <= DS_ACT_SOFTPLUS, "%s: unknown activation id %d", who, act);
  DS_CHECK(mode >= DS_MODE_FP32 && mode <= DS_MODE_TF32X3, "%s: unknown mode %d", who, mode);
  return 0;
}

// T_1..T_{K-1} into basis ([K-1, B, M, Fin]); T_0 = x
int compute_basis(const ds_plan* plan, int32_t recursion, int32_t K, int64_t B, int64_t Fin, const float* x,
                  float* basis, cudaStream_t st, bool transpose = false) {
  const SparseDev& S = transpose ? plan->bwd : plan->fwd;  // the lattice path requires L~ symmetric
  const int64_t A = B * S.M * Fin;
  if (lattice_usable(plan, K, B, Fin)) {
    // the hops run fused on chip (ds_lattice.cu), every hop's own-tile result goes to its basis slot.  The lattice
    // tables carry an H-ring halo: K - 1 <= H hops are one pass; longer recursions (the reference ships K = 10,
    // examples/quick_start.ipynb:118-127) are CHAINED in passes of <= H hops - pass j starts from cur = T_{jH} and, for
    // Chebyshev, takes old = T_{jH-1} as the first hop's additive input (2 L~ T_{jH} - T_{jH-1}); HBM sees each basis
    // tensor written once and the two seam tensors read once more, instead of 3 tensors per hop
    const int H = lattice_halo(plan);
    auto T = [&](int k) -> const float* { return k == 0 ? x : basis + (int64_t)(k - 1) * A; };
    for (int k0 = 0; k0 < K - 1; k0 += H) {
      const int n = std::min(H, K - 1 - k0);
      const float* add[16] = {};
      float* out[16] = {};
      float al[16], be[16], ga[16];
      for (int s = 1; s <= n; ++s) {
        const int k = k0 + s;  // this hop produces T_k
        const bool cheb2 = recursion == DS_RECURSION_CHEBYSHEV && k >= 2;
        al[s - 1] = cheb2 ? 2.f : 1.f;
        be[s - 1] = cheb2 && s >= 2 ? -1.f : 0.f;   // old = the pass's previous tensor (on chip from the 2nd hop on)
        ga[s - 1] = 0.f;
        if (cheb2 && s == 1) {                       // seam: T_{k-2} comes from HBM
          add[0] = T(k - 2);
          ga[0] = -1.f;
        }
        out[s - 1] = basis + (int64_t)(k - 1) * A;
      }
      DS_TRY(lattice_recursion(plan, B, (int)Fin, n, T(k0), add, out, al, be, ga, st));
    }
    return 0;
  }
  auto T = [&](int k) -> const float* { return k == 0 ? x : basis + (int64_t)(k - 1) * A; };
  for (int k = 1; k < K; ++k) {
    float* out = basis + (int64_t)(k - 1) * A;
    if (recursion == DS_RECURSION_CHEBYSHEV && k >= 2) {
      DS_TRY(launch_spmm(S, B, Fin, T(k - 1), 2.f, T(k - 2), -1.f, nullptr, 0.f, out, st));  // gnn_layers.py:141
    } else {
      DS_TRY(launch_spmm(S, B, Fin, T(k - 1), 1.f, nullptr, 0.f, nullptr, 0.f, out, st));  // :138 / :288
    }
  }
  return 0;
}

}  // namespace
}  // namespace ds

extern "C" {

int64_t ds_graph_conv_basis_elems(int64_t M, int64_t B, int64_t Fin, int32_t K) {
  return K > 1 ? (int64_t)(K - 1) * B * M * Fin : 0;
}

int32_t ds_graph_conv_forward_writes_basis(const ds_plan_t* plan, int32_t K, int64_t B, int64_t Fin, int64_t Fout,
                                           int32_t mode) {
  using namespace ds;
  if (plan == nullptr || K <= 1) return 0;
  return fused_conv_usable(plan, K, B, Fin, Fout, mode) ? 0 : 1;
}

int ds_graph_conv_basis(const ds_plan_t* plan, int32_t recursion, int32_t K, int64_t B, int64_t F, const float* x,
                        float* basis, int32_t transpose, void* stream) {
  using namespace ds;
  DS_TRY(check_common(plan, recursion, K, B, F, F, DS_ACT_LINEAR, DS_MODE_FP32, "ds_graph_conv_basis"));
  DS_CHECK(x && (K == 1 || basis), "ds_graph_conv_basis: NULL tensor");
  return compute_basis(plan, recursion, K, B, F, x, basis, (cudaStream_t)stream, transpose != 0);
}

int ds_graph_conv_forward(const ds_plan_t* plan, int32_t recursion, int32_t K, int64_t B, int64_t Fin, int64_t Fout,
                          const float* x, const float* kernel, const float* bias, int32_t act, float* y, float* basis,
                          int32_t mode, void* stream) {
  using namespace ds;
  DS_TRY(check_common(plan, recursion, K, B, Fin, Fout, act, mode, "ds_graph_conv_forward"));
  DS_CHECK(x && kernel && y, "ds_graph_conv_forward: NULL tensor");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t M = plan->M, R = B * M, A = R * Fin;
  const bool fused_fwd = fused_conv_usable(plan, K, B, Fin, Fout, mode);
  DS_CHECK(K == 1 || fused_fwd || basis != nullptr,
           "ds_graph_conv_forward: basis workspace required for K > 1 (see ds_graph_conv_forward_writes_basis)");
  if (fused_fwd) {
    // recursion + contraction in one kernel; the basis never leaves the chip (and `basis` is left untouched:
    // the fused backward does not need it)
    return fused_conv(plan, recursion, K, B, Fin, Fout, x, nullptr, kernel, (int64_t)K * Fout, Fout, 1, bias, act, y,
                      mode, st);
  }
  DS_TRY(compute_basis(plan, recursion, K, B, Fin, x, basis, st));
  if (mode == DS_MODE_FP32) {
    return launch_gemm_nn(R, Fout, Fin, K, x, basis, A, Fin, kernel, Fout, K, 1, bias, Fout, act, y, Fout, st);
  }
  DS_CHECK(umma_supported(Fin, K, Fout) == 0,
           "ds_graph_conv_forward: tensor-core mode needs Fin %% 8 == 0 and Fout %% 16 == 0, 16 <= Fout <= 256 "
           "(got Fin=%lld Fout=%lld); use DS_MODE_FP32",
           (long long)Fin, (long long)Fout);
  return launch_umma_gemm_nn(R, Fout, Fin, K, x, basis, A, kernel, K, 1, bias, act, y, mode, st);
}

int64_t ds_graph_conv_backward_workspace_elems(int64_t M, int64_t B, int64_t Fin, int64_t Fout, int32_t K,
                                               int32_t have_basis, int32_t act) {
  using namespace ds;
  const int64_t R = B * M, A = R * Fin;
  int64_t n = 0;
  if (act != DS_ACT_LINEAR) n += R * Fout;                  // dz
  if (!have_basis && K > 1) n += (int64_t)(K - 1) * A;      // recomputed basis
  if (K > 1) n += (int64_t)(K - 1) * R * Fout;              // T_1..T_{K-1} of L~^T applied to dz
  n += umma_tn_workspace_elems(R, Fin, Fout, K);            // fused path: dkernel partials in the transposed form
  n += std::max(gemm_tn_workspace_elems(R, Fin, K, Fout),   // dkernel split partials (fp32 / tensor-core kernel)
                umma_tn_workspace_elems(R, Fout, Fin, K));
  n += colsum_workspace_elems(Fout);                        // dbias partials
  return n + 64;
}

int ds_graph_conv_backward(const ds_plan_t* plan, int32_t recursion, int32_t K, int64_t B, int64_t Fin, int64_t Fout,
                           const float* x, const float* kernel, const float* y, const float* dy, int32_t act,
                           const float* basis, float* dx, float* dkernel, float* dbias, float* workspace, int32_t mode,
                           void* stream) {
  using namespace ds;
  DS_TRY(check_common(plan, recursion, K, B, Fin, Fout, act, mode, "ds_graph_conv_backward"));
  DS_CHECK(x && kernel && dy && dkernel && workspace, "ds_graph_conv_backward: NULL tensor");
  DS_CHECK(act == DS_ACT_LINEAR || y != nullptr, "ds_graph_conv_backward: y required when act != LINEAR");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t M = plan->M, R = B * M, A = R * Fin;
  float* ws = workspace;
  auto take = [&](int64_t n) { float* p = ws; ws += (n + 3) / 4 * 4; return p; };

  // 1. gradient w.r.t. the pre-activation (and, when the opt-in streaming kernel applies, dbias in the same sweep)
  const float* dz = dy;
  float* cs_partial = take(colsum_workspace_elems(Fout));
  bool dbias_done = dbias == nullptr;
  if (act != DS_ACT_LINEAR) {
    float* dzb = take(R * Fout);
    const int rc = launch_act_backward_colsum(R, Fout, y, dy, act, dzb, dbias, cs_partial, st);
    if (rc > 0) return rc;
    if (rc == 0) dbias_done = true;
    else DS_TRY(launch_act_backward(R, Fout, Fout, y, dy, act, dzb, st));
    dz = dzb;
  }
  // fused path: dx = fused conv on dz (also emits U_k = T_k(L~^T) dz), dkernel = [dz|U_1..]^T-contracted with x
  if (dx != nullptr && K > 1 && plan->symmetric && fused_conv_usable(plan, K, B, Fout, Fin, mode) &&
      umma_tn_supported(R, Fin, Fout, K) == 0) {
    float* U = take((int64_t)(K - 1) * R * Fout);
    float* tn_partial = take(umma_tn_workspace_elems(R, Fin, Fout, K));
    if (!dbias_done) DS_TRY(launch_colsum(R, Fout, Fout, dz, dbias, cs_partial, st));
    // B_k(o, f) = kernel[(f*K + k)*Fout + o]
    DS_TRY(fused_conv(plan, recursion, K, B, Fout, Fin, dz, U, kernel, 1, Fout, (int64_t)K * Fout, nullptr,
                      DS_ACT_LINEAR, dx, mode, st));
    // dkernel[(f*K + k)*Fout + o] = sum_r U_k[r, o] x[r, f]   (adjoint identity: sum_m T_k(L~)x . dz = sum_m x . T_k(L~^T)dz)
    return launch_umma_gemm_tn(R, Fin, Fout, K, dz, U, x, dkernel, 1, Fout, (int64_t)K * Fout, tn_partial, mode, st);
  }
  // 2. basis (saved by the forward or recomputed)
  const float* T = basis;
  if (T == nullptr && K > 1) {
    float* tb = take((int64_t)(K - 1) * A);
    DS_TRY(compute_basis(plan, recursion, K, B, Fin, x, tb, st));
    T = tb;
  }
  float* U = K > 1 ? take((int64_t)(K - 1) * R * Fout) : nullptr;
  float* tn_partial = take(std::max(gemm_tn_workspace_elems(R, Fin, K, Fout), umma_tn_workspace_elems(R, Fout, Fin, K)));

  // 3. dbias = sum_{b,m} dz
  if (!dbias_done) DS_TRY(launch_colsum(R, Fout, Fout, dz, dbias, cs_partial, st));
  // 4. dkernel[f*K + k, o] = sum_{b,m} T_k[b,m,f] dz[b,m,o]
  if (mode != DS_MODE_FP32 && umma_tn_supported(R, Fout, Fin, K) == 0) {
    DS_TRY(launch_umma_gemm_tn(R, Fout, Fin, K, x, T, dz, dkernel, (int64_t)K * Fout, Fout, 1, tn_partial, mode, st));
  } else {
    DS_TRY(launch_gemm_tn(R, Fout, Fin, K, x, T, A, Fin, dz, Fout, dkernel, Fout, K, 1, tn_partial, st));
  }
  if (dx == nullptr) return 0;

  // 5. dx = sum_k T_k(L~^T)(dz) W_k^T.  L~ acts on the pixel axis and W_k on the channel axis, so they
  //    commute: the data gradient is the FORWARD pipeline applied to dz - the same (fused) recursion with
  //    L~^T on Fout channels, then one segmented contraction with B(k, o, f) = kernel[(f*K + k), o].
  //    (Equivalent to the Clenshaw adjoint sum of SURVEY a18, with K-1 hops and a single GEMM.)
  DS_TRY(compute_basis(plan, recursion, K, B, Fout, dz, U, st, /*transpose=*/true));
  if (mode != DS_MODE_FP32 && umma_supported(Fout, K, Fin) == 0)
    return launch_umma_gemm(R, Fin, Fout, K, dz, U, R, kernel, /*b_k_stride=*/1, /*b_seg_stride=*/Fout,
                            /*b_n_stride=*/(int64_t)K * Fout, nullptr, 1, DS_ACT_LINEAR, dx, Fin, mode, st);
  return launch_gemm_nt_seg(R, Fin, Fout, K, dz, U, R * Fout, Fout, kernel, Fout, K, 1, dx, Fin, st);
}

int ds_bias_act_forward(int64_t R, int64_t F, const float* z, const float* bias, int32_t act, float* y, void* stream);
int ds_bias_act_backward(int64_t R, int64_t F, const float* y, const float* dy, int32_t act, float* dz, float* dbias,
                         float* workspace, void* stream) {
  using namespace ds;
  DS_CHECK(dy && dz && R > 0 && F > 0, "ds_bias_act_backward: bad argument");
  DS_CHECK(act == DS_ACT_LINEAR || y != nullptr, "ds_bias_act_backward: y required");
  cudaStream_t st = (cudaStream_t)stream;
  if (act != DS_ACT_LINEAR) {
    const int rc = workspace != nullptr ? launch_act_backward_colsum(R, F, y, dy, act, dz, dbias, workspace, st) : -1;
    if (rc >= 0) return rc;  // dz and dbias in one sweep (opt-in streaming kernel)
    DS_TRY(launch_act_backward(R, F, F, y, dy, act, dz, st));
  } else if (dz != dy) {
    DS_CUDA(cudaMemcpyAsync(dz, dy, sizeof(float) * R * F, cudaMemcpyDeviceToDevice, st));
  }
  if (dbias != nullptr) {
    DS_CHECK(workspace != nullptr, "ds_bias_act_backward: workspace required for dbias");
    DS_TRY(launch_colsum(R, F, F, dz, dbias, workspace, st));
  }
  return 0;
}

// ---- pseudo convolutions (healpy_layers.py:87-216) as GEMMs over NESTED-contiguous children ----

static int pconv_check(int64_t B, int64_t M, int64_t Fin, int64_t Fout, int32_t p, int32_t act, int32_t mode,
                       bool reduce, const char* who) {
  using namespace ds;
  DS_CHECK(p >= 1 && p <= 12, "%s: p=%d out of range", who, p);
  DS_CHECK(B >= 1 && M >= 1 && Fin >= 1 && Fout >= 1, "%s: empty shape", who);
  DS_CHECK(!reduce || M % (1LL << (2 * p)) == 0, "%s: M=%lld not divisible by 4^p=%lld", who, (long long)M,
           (long long)(1LL << (2 * p)));
  DS_CHECK(act >= DS_ACT_LINEAR && act <= DS_ACT_SOFTPLUS, "%s: unknown activation id %d", who, act);
  DS_CHECK(mode >= DS_MODE_FP32 && mode <= DS_MODE_TF32X3, "%s: unknown mode %d", who, mode);
  return 0;
}

int ds_pconv_forward(int64_t B, int64_t M, int64_t Fin, int64_t Fout, int32_t p, const float* x, const float* w,
                     const float* bias, int32_t act, float* y, int32_t mode, void* stream) {
  using namespace ds;
  DS_TRY(pconv_check(B, M, Fin, Fout, p, act, mode, true, "ds_pconv_forward"));
  DS_CHECK(x && w && y, "ds_pconv_forward: NULL tensor");
  const int64_t r = 1LL << (2 * p), Ro = B * (M / r), Kc = r * Fin;
  if (skinny_usable(Ro, Kc, Fout)) {  // short reduction (e.g. 4 -> 16 at the head of a network): streaming kernel
    const int rc = launch_skinny_nn(Ro, Fout, Kc, x, w, bias, Fout, act, y, (cudaStream_t)stream);
    if (rc >= 0) return rc;
  }
  return launch_gemm_nn(Ro, Fout, Kc, 1, x, nullptr, 0, Kc, w, Fout, 1, 0, bias, Fout, act, y, Fout,
                        (cudaStream_t)stream);
}

int64_t ds_pconv_backward_workspace_elems(int64_t B, int64_t M, int64_t Fin, int64_t Fout, int32_t p, int32_t act) {
  using namespace ds;
  const int64_t r = 1LL << (2 * p), Ro = B * (M / r), Kc = r * Fin;
  return (act != DS_ACT_LINEAR ? Ro * Fout : 0) + gemm_tn_workspace_elems(Ro, Kc, 1, Fout) +
         colsum_workspace_elems(Fout) + skinny_bwd_workspace_elems(Ro, Kc, Fout) + 64;
}

int ds_pconv_backward(int64_t B, int64_t M, int64_t Fin, int64_t Fout, int32_t p, const float* x, const float* w,
                      const float* y, const float* dy, int32_t act, float* dx, float* dw, float* dbias,
                      float* workspace, int32_t mode, void* stream) {
  using namespace ds;
  DS_TRY(pconv_check(B, M, Fin, Fout, p, act, mode, true, "ds_pconv_backward"));
  DS_CHECK(x && w && dy && dw && workspace, "ds_pconv_backward: NULL tensor");
  DS_CHECK(act == DS_ACT_LINEAR || y != nullptr, "ds_pconv_backward: y required when act != LINEAR");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t r = 1LL << (2 * p), Ro = B * (M / r), Kc = r * Fin;
  float* ws = workspace;
  auto take = [&](int64_t n) { float* q = ws; ws += (n + 3) / 4 * 4; return q; };
  float* dzb = act != DS_ACT_LINEAR ? take(Ro * Fout) : nullptr;
  float* tn_partial = take(gemm_tn_workspace_elems(Ro, Kc, 1, Fout));
  float* cs_partial = take(colsum_workspace_elems(Fout));
  const float* dz = dy;
  bool done = false;  // dz, dw, dbias produced by the fused streaming sweep
  if (skinny_usable(Ro, Kc, Fout)) {
    float* sk_partial = take(skinny_bwd_workspace_elems(Ro, Kc, Fout));
    const int rc = launch_skinny_pconv_bwd(Ro, Fout, Kc, x, y, dy, act, dx ? dzb : nullptr, dw, dbias, sk_partial, st);
    if (rc > 0) return rc;
    done = rc == 0;
    if (done && dzb != nullptr) dz = dzb;  // only valid (and only needed) when dx is wanted
  }
  if (!done) {
    if (act != DS_ACT_LINEAR) {
      DS_TRY(launch_act_backward(Ro, Fout, Fout, y, dy, act, dzb, st));
      dz = dzb;
    }
    if (dbias) DS_TRY(launch_colsum(Ro, Fout, Fout, dz, dbias, cs_partial, st));
    DS_TRY(launch_gemm_tn(Ro, Fout, Kc, 1, x, nullptr, 0, Kc, dz, Fout, dw, Fout, 1, 0, tn_partial, st));
  }
  if (dx) DS_TRY(launch_gemm_nt(Ro, Kc, Fout, 1, dz, Fout, w, Fout, 1, 0, nullptr, 1, DS_ACT_LINEAR, dx, Kc, 0, st));
  return 0;
}

int ds_pconvT_forward(int64_t B, int64_t M, int64_t Fin, int64_t Fout, int32_t p, const float* x, const float* w,
                      const float* bias, int32_t act, float* y, int32_t mode, void* stream) {
  using namespace ds;
  DS_TRY(pconv_check(B, M, Fin, Fout, p, act, mode, false, "ds_pconvT_forward"));
  DS_CHECK(x && w && y, "ds_pconvT_forward: NULL tensor");
  const int64_t r = 1LL << (2 * p), R = B * M, Nc = r * Fout;
  return launch_gemm_nt(R, Nc, Fin, 1, x, Fin, w, Fin, 1, 0, bias, Fout, act, y, Nc, 0, (cudaStream_t)stream);
}

int64_t ds_pconvT_backward_workspace_elems(int64_t B, int64_t M, int64_t Fin, int64_t Fout, int32_t p, int32_t act) {
  using namespace ds;
  const int64_t r = 1LL << (2 * p), R = B * M, Nc = r * Fout;
  return (act != DS_ACT_LINEAR ? R * Nc : 0) + gemm_tn_workspace_elems(R, Nc, 1, Fin) + colsum_workspace_elems(Nc) +
         64;
}

int ds_pconvT_backward(int64_t B, int64_t M, int64_t Fin, int64_t Fout, int32_t p, const float* x, const float* w,
                       const float* y, const float* dy, int32_t act, float* dx, float* dw, float* dbias,
                       float* workspace, int32_t mode, void* stream) {
  using namespace ds;
  DS_TRY(pconv_check(B, M, Fin, Fout, p, act, mode, false, "ds_pconvT_backward"));
  DS_CHECK(x && w && dy && dw && workspace, "ds_pconvT_backward: NULL tensor");
  DS_CHECK(act == DS_ACT_LINEAR || y != nullptr, "ds_pconvT_backward: y required when act != LINEAR");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t r = 1LL << (2 * p), R = B * M, Nc = r * Fout;
  float* ws = workspace;
  auto take = [&](int64_t n) { float* q = ws; ws += (n + 3) / 4 * 4; return q; };
  const float* dz = dy;
  if (act != DS_ACT_LINEAR) {
    float* dzb = take(R * Nc);
    DS_TRY(launch_act_backward(R, Nc, Fout, y, dy, act, dzb, st));
    dz = dzb;
  }
  float* tn_partial = take(gemm_tn_workspace_elems(R, Nc, 1, Fin));
  float* cs_partial = take(colsum_workspace_elems(Nc));
  if (dbias) DS_TRY(launch_colsum(R, Nc, Fout, dz, dbias, cs_partial, st));
  // dw[c*Fout+o, f] = sum_r dz[r, c*Fout+o] x[r, f]
  DS_TRY(launch_gemm_tn(R, Fin, Nc, 1, dz, nullptr, 0, Nc, x, Fin, dw, Fin, 1, 0, tn_partial, st));
  // dx[r, f] = sum_{c,o} dz[r, c*Fout+o] w[c*Fout+o, f]
  if (dx) DS_TRY(launch_gemm_nn(R, Fin, Nc, 1, dz, nullptr, 0, Nc, w, Fin, 1, 0, nullptr, 1, DS_ACT_LINEAR, dx, Fin, st));
  return 0;
}

}  // extern "C"
