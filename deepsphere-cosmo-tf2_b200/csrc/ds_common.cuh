// Shared helpers for the deepsphere_b200 C-ABI library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>

#include "../../include/deepsphere_b200.h"

namespace ds {

// ---- error plumbing ---------------------------------------------------------------
std::string& last_error_ref();
int fail(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

#define DS_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return ds::fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define DS_CHECK(cond, ...)                 \
  do {                                      \
    if (!(cond)) return ds::fail(__VA_ARGS__); \
  } while (0)

#define DS_TRY(expr)        \
  do {                      \
    int _rc = (expr);       \
    if (_rc != 0) return _rc; \
  } while (0)

// after a kernel launch
#define DS_LAUNCHED()                                  \
  do {                                                 \
    ds::g_launches.fetch_add(1, std::memory_order_relaxed); \
    DS_CUDA(cudaGetLastError());                       \
  } while (0)

// Function attributes (opt-in dynamic shared memory, carve-out) belong to the device a kernel is loaded on, so a
// process that drives several GPUs has to set them once per device, not once per process.
//   static PerDeviceOnce once;  DS_TRY(once.run([&]() -> int { DS_CUDA(cudaFuncSetAttribute(...)); return 0; }));
struct PerDeviceOnce {
  std::mutex m;
  uint64_t done[4] = {0, 0, 0, 0};  // bit per device ordinal (256 devices)
  template <class F>
  int run(F&& f) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 256) dev = 0;
    std::lock_guard<std::mutex> lock(m);
    const uint64_t bit = 1ull << (dev & 63);
    if (done[dev >> 6] & bit) return 0;
    const int rc = f();
    if (rc == 0) done[dev >> 6] |= bit;
    return rc;
  }
};

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// ---- plan (device-resident ELL + CSR tail of L~ and of L~^T) -------------------------
struct SparseDev {
  int64_t M = 0;
  int32_t W = 0;              // ELL width
  int32_t* ell_col = nullptr; // [M, W]  padded with col = row, val = 0
  float* ell_val = nullptr;   // [M, W]
  int32_t Wp = 0;             // packed row length (W rounded up to even)
  int4* ell_pk = nullptr;     // [M, Wp/2] of (col0, val0 bits, col1, val1 bits): one 16-byte load = 2 entries
  int64_t n_tail_rows = 0;    // rows with more than W entries
  int64_t tail_nnz = 0;
  int32_t* tail_rows = nullptr;   // [n_tail_rows]
  int64_t* tail_rowptr = nullptr; // [n_tail_rows + 1]
  int32_t* tail_col = nullptr;    // [tail_nnz]
  float* tail_val = nullptr;      // [tail_nnz]
};

}  // namespace ds

namespace ds { struct LatticeAttachment; }

struct ds_plan {
  ds::LatticeAttachment* lattice = nullptr;  // optional fused-recursion plan (ds_lattice_api.cu)
  int64_t M = 0;
  int64_t nnz = 0;
  int device = 0;
  bool symmetric = false;
  ds::SparseDev fwd;  // L~
  ds::SparseDev bwd;  // L~^T (aliases fwd when symmetric)
  int64_t device_bytes = 0;
};

namespace ds {

// ---- activations (tf.keras.activations semantics) ------------------------------------
__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case DS_ACT_RELU: return fmaxf(v, 0.f);
    case DS_ACT_ELU: return v > 0.f ? v : expm1f(v);
    case DS_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case DS_ACT_TANH: return tanhf(v);
    case DS_ACT_SOFTPLUS: return v > 20.f ? v : log1pf(expf(v));
    default: return v;
  }
}
// derivative expressed through the OUTPUT y = act(z)
__device__ __forceinline__ float act_grad_from_y(float y, int act) {
  switch (act) {
    case DS_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case DS_ACT_ELU: return y > 0.f ? 1.f : y + 1.f;
    case DS_ACT_SIGMOID: return y * (1.f - y);
    case DS_ACT_TANH: return 1.f - y * y;
    case DS_ACT_SOFTPLUS: return 1.f - expf(-y);
    default: return 1.f;
  }
}

// ---- internal kernels' host launchers (defined in the .cu files) ----------------------
// out = alpha * S in + beta * prev + gamma * add
int launch_spmm(const SparseDev& S, int64_t B, int64_t F, const float* in, float alpha, const float* prev, float beta,
                const float* add, float gamma, float* out, cudaStream_t st);

// C[R,N] = act( sum_seg A_seg[R,Kc] * Bm[kc*b_kc_stride + seg*b_seg_stride, :N] + bias[col % bias_mod] )
//   A_seg = (seg == 0 ? A0 : Arest + (seg-1)*a_seg_stride), row stride lda; Bm row stride ldb; C row stride ldc
int launch_gemm_nn(int64_t R, int64_t N, int64_t Kc, int nseg, const float* A0, const float* Arest,
                   int64_t a_seg_stride, int64_t lda, const float* Bm, int64_t ldb, int64_t b_kc_stride,
                   int64_t b_seg_stride, const float* bias, int64_t bias_mod, int act, float* C, int64_t ldc,
                   cudaStream_t st);
// C_seg[R,Nc] = act( A[R,Kd] * Bm[nc*b_nc_stride + seg*b_seg_stride, :Kd]^T + bias[(col) % bias_mod] ), seg < nseg
//   C_seg = C + seg*c_seg_stride
int launch_gemm_nt(int64_t R, int64_t Nc, int64_t Kd, int nseg, const float* A, int64_t lda, const float* Bm,
                   int64_t ldb, int64_t b_nc_stride, int64_t b_seg_stride, const float* bias, int64_t bias_mod,
                   int act, float* C, int64_t ldc, int64_t c_seg_stride, cudaStream_t st);
// C[R,Nc] = sum_seg A_seg[R,Kd] * Bm[nc*b_nc_stride + seg*b_seg_stride, :Kd]^T
int launch_gemm_nt_seg(int64_t R, int64_t Nc, int64_t Kd, int nseg, const float* A0, const float* Arest,
                       int64_t a_seg_stride, int64_t lda, const float* Bm, int64_t ldb, int64_t b_nc_stride,
                       int64_t b_seg_stride, float* C, int64_t ldc, cudaStream_t st);
// C[kc*c_kc_stride + seg*c_seg_stride, :N] = sum_r A_seg[r,kc] * D[r,:N]     (reduction over R rows)
//   partial: workspace of gemm_tn_workspace_elems floats
int64_t gemm_tn_workspace_elems(int64_t R, int64_t Kc, int nseg, int64_t N);
int launch_gemm_tn(int64_t R, int64_t N, int64_t Kc, int nseg, const float* A0, const float* Arest,
                   int64_t a_seg_stride, int64_t lda, const float* D, int64_t ldd, float* C, int64_t ldc,
                   int64_t c_kc_stride, int64_t c_seg_stride, float* partial, cudaStream_t st);
// dz = dy * act'(y);  optional column sums of dz into dbias[F] (bias_mod folding: col % F)
int launch_act_backward(int64_t R, int64_t Ncols, int64_t F, const float* y, const float* dy, int act, float* dz,
                        cudaStream_t st);
// colsum[c % F] = sum over rows and folded columns of Z[R, Ncols];  workspace colsum_workspace_elems floats
int64_t colsum_workspace_elems(int64_t Ncols);
int launch_colsum(int64_t R, int64_t Ncols, int64_t F, const float* Z, float* out, float* workspace, cudaStream_t st);
int launch_colsum_final(int64_t Ncols, int64_t F, int nblk, const float* partial, float* out, cudaStream_t st);

// streaming kernels for narrow contractions (ds_narrow.cu: few output columns, short reduction per segment - the 1 - 5
// channel layers of the reference's notebooks); B(seg, k, n) = Bm[k * sk + seg * ss + n * sn]
bool narrow_rows_usable(int64_t N, int64_t K, int nseg);
bool narrow_tn_usable(int64_t N, int64_t K, int nseg);
int64_t narrow_tn_workspace_elems(int64_t R, int64_t K, int nseg, int64_t N);
int launch_narrow_rows(int64_t R, int64_t N, int64_t K, int nseg, const float* A0, const float* Arest, int64_t a_seg_stride,
                       int64_t lda, const float* Bm, int64_t sk, int64_t ss, int64_t sn, const float* bias, int64_t bias_mod,
                       int act, float* C, int64_t ldc, cudaStream_t st);
int launch_narrow_tn(int64_t R, int64_t N, int64_t K, int nseg, const float* A0, const float* Arest, int64_t a_seg_stride,
                     int64_t lda, const float* D, int64_t ldd, float* partial, int64_t* nblk, cudaStream_t st);

// streaming kernels for contractions with a very short reduction (ds_skinny.cu; opt-in with DEEPSPHERE_SKINNY=1).
// The launchers return -1 (nothing launched) when an operand is not 16-byte aligned: the caller then takes the
// tiled kernels; 0 = done, > 0 = error (ds_last_error).
bool skinny_usable(int64_t R, int64_t Kc, int64_t N);
int64_t skinny_bwd_workspace_elems(int64_t R, int64_t Kc, int64_t N);
int launch_skinny_nn(int64_t R, int64_t N, int64_t Kc, const float* A, const float* Bm, const float* bias,
                     int64_t bias_mod, int act, float* C, cudaStream_t st);
// dz = dy * act'(y) and dbias = colsum(dz) in one sweep; workspace: colsum_workspace_elems(F) floats
int launch_act_backward_colsum(int64_t R, int64_t F, const float* y, const float* dy, int act, float* dz, float* dbias,
                               float* workspace, cudaStream_t st);
int launch_skinny_pconv_bwd(int64_t R, int64_t N, int64_t Kc, const float* X, const float* y, const float* dy, int act,
                            float* dz_out, float* dw, float* dbias, float* partial, cudaStream_t st);

}  // namespace ds
