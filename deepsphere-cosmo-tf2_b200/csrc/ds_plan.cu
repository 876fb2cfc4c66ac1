// Plan builder: host COO (the reference's _L_indices/_L_values, gnn_layers.py:68-72) ->
// row-major sorted (tf.sparse.reorder, gnn_layers.py:115) -> fixed-width ELL + CSR tail
// for L~ and L~^T, uploaded to the current device.  Also: error plumbing, misc entries.
#include <algorithm>
#include <cstdarg>
#include <numeric>
#include <vector>

#include "ds_common.cuh"

namespace ds {

void lattice_free(LatticeAttachment* L);  // ds_lattice_api.cu

std::string& last_error_ref() {
  thread_local std::string s;
  return s;
}
int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return 1;
}
std::atomic<int64_t> g_launches{0};

namespace {

struct HostCsr {
  int64_t M = 0;
  std::vector<int64_t> rowptr;
  std::vector<int32_t> col;
  std::vector<float> val;
};

// stable counting sort by row, then sort each row by column
void coo_to_csr(int64_t M, int64_t nnz, const int64_t* rows, int64_t rstride, const int64_t* cols, int64_t cstride,
                const float* values, HostCsr& out) {
  out.M = M;
  out.rowptr.assign(M + 1, 0);
  for (int64_t i = 0; i < nnz; ++i) out.rowptr[rows[i * rstride] + 1]++;
  for (int64_t r = 0; r < M; ++r) out.rowptr[r + 1] += out.rowptr[r];
  out.col.resize(nnz);
  out.val.resize(nnz);
  std::vector<int64_t> fill(out.rowptr.begin(), out.rowptr.end() - 1);
  for (int64_t i = 0; i < nnz; ++i) {
    int64_t p = fill[rows[i * rstride]]++;
    out.col[p] = (int32_t)cols[i * cstride];
    out.val[p] = values[i];
  }
  std::vector<std::pair<int32_t, float>> tmp;
  for (int64_t r = 0; r < M; ++r) {
    int64_t a = out.rowptr[r], b = out.rowptr[r + 1];
    bool sorted = true;
    for (int64_t p = a + 1; p < b; ++p)
      if (out.col[p] < out.col[p - 1]) { sorted = false; break; }
    if (sorted) continue;
    tmp.resize(b - a);
    for (int64_t p = a; p < b; ++p) tmp[p - a] = {out.col[p], out.val[p]};
    std::stable_sort(tmp.begin(), tmp.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
    for (int64_t p = a; p < b; ++p) { out.col[p] = tmp[p - a].first; out.val[p] = tmp[p - a].second; }
  }
}

int32_t choose_width(const HostCsr& A) {
  if (A.M == 0) return 1;
  std::vector<int32_t> len(A.M);
  for (int64_t r = 0; r < A.M; ++r) len[r] = (int32_t)(A.rowptr[r + 1] - A.rowptr[r]);
  std::vector<int32_t> s(len);
  std::sort(s.begin(), s.end());
  int32_t mx = s.back();
  int32_t p995 = s[(size_t)std::min<int64_t>(A.M - 1, (int64_t)(0.995 * (double)A.M))];
  int32_t w = (mx <= p995 + p995 / 4 + 1) ? mx : p995;
  return std::max(w, 1);
}

template <typename T>
int upload(const std::vector<T>& h, T** d, int64_t& bytes) {
  *d = nullptr;
  size_t n = std::max<size_t>(h.size(), 1) * sizeof(T);
  DS_CUDA(cudaMalloc((void**)d, n));
  if (!h.empty()) DS_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  bytes += (int64_t)n;
  return 0;
}

int build_sparse_dev(const HostCsr& A, int32_t width, SparseDev& S, int64_t& bytes) {
  const int64_t M = A.M;
  const int32_t W = width > 0 ? width : choose_width(A);
  S.M = M;
  S.W = W;
  std::vector<int32_t> ecol((size_t)M * W);
  std::vector<float> eval((size_t)M * W, 0.f);
  std::vector<int32_t> trows;
  std::vector<int64_t> trowptr(1, 0);
  std::vector<int32_t> tcol;
  std::vector<float> tval;
  for (int64_t r = 0; r < M; ++r) {
    int64_t a = A.rowptr[r], b = A.rowptr[r + 1];
    int64_t n = b - a;
    for (int32_t j = 0; j < W; ++j) {
      if (j < n) {
        ecol[(size_t)r * W + j] = A.col[a + j];
        eval[(size_t)r * W + j] = A.val[a + j];
      } else {
        ecol[(size_t)r * W + j] = (int32_t)r;
      }
    }
    if (n > W) {
      trows.push_back((int32_t)r);
      for (int64_t p = a + W; p < b; ++p) { tcol.push_back(A.col[p]); tval.push_back(A.val[p]); }
      trowptr.push_back((int64_t)tcol.size());
    }
  }
  // packed copy for the vectorised hop kernel: (col, val) pairs, two per 16-byte word
  S.Wp = (W + 1) & ~1;
  std::vector<int4> pk((size_t)M * (S.Wp / 2));
  for (int64_t r = 0; r < M; ++r)
    for (int32_t j = 0; j < S.Wp; j += 2) {
      int4 q;
      q.x = j < W ? ecol[(size_t)r * W + j] : (int32_t)r;
      const float v0 = j < W ? eval[(size_t)r * W + j] : 0.f;
      q.z = j + 1 < W ? ecol[(size_t)r * W + j + 1] : (int32_t)r;
      const float v1 = j + 1 < W ? eval[(size_t)r * W + j + 1] : 0.f;
      std::memcpy(&q.y, &v0, 4);
      std::memcpy(&q.w, &v1, 4);
      pk[(size_t)r * (S.Wp / 2) + j / 2] = q;
    }
  DS_TRY(upload(pk, &S.ell_pk, bytes));
  S.n_tail_rows = (int64_t)trows.size();
  S.tail_nnz = (int64_t)tcol.size();
  DS_TRY(upload(ecol, &S.ell_col, bytes));
  DS_TRY(upload(eval, &S.ell_val, bytes));
  DS_TRY(upload(trows, &S.tail_rows, bytes));
  DS_TRY(upload(trowptr, &S.tail_rowptr, bytes));
  DS_TRY(upload(tcol, &S.tail_col, bytes));
  DS_TRY(upload(tval, &S.tail_val, bytes));
  return 0;
}

void free_sparse_dev(SparseDev& S) {
  cudaFree(S.ell_col);
  cudaFree(S.ell_val);
  cudaFree(S.ell_pk);
  cudaFree(S.tail_rows);
  cudaFree(S.tail_rowptr);
  cudaFree(S.tail_col);
  cudaFree(S.tail_val);
  S = SparseDev();
}

}  // namespace
}  // namespace ds

extern "C" {

int ds_abi_version(void) { return DS_ABI_VERSION; }

const char* ds_last_error(void) { return ds::last_error_ref().c_str(); }

int ds_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int64_t ds_launch_count(void) { return ds::g_launches.load(); }

int ds_plan_create_coo(int64_t M, int64_t nnz, const int64_t* indices, const float* values, int32_t ell_width,
                       ds_plan_t** plan_out) {
  using namespace ds;
  DS_CHECK(plan_out != nullptr, "ds_plan_create_coo: plan_out is NULL");
  *plan_out = nullptr;
  DS_CHECK(M > 0 && M < (int64_t)INT32_MAX, "ds_plan_create_coo: M=%lld out of range", (long long)M);
  DS_CHECK(nnz >= 0, "ds_plan_create_coo: negative nnz");
  DS_CHECK(nnz == 0 || (indices && values), "ds_plan_create_coo: NULL indices/values");
  DS_CHECK(ds_device_count() > 0, "ds_plan_create_coo: no CUDA device (there is no CPU fallback)");
  for (int64_t i = 0; i < nnz; ++i) {
    int64_t r = indices[2 * i], c = indices[2 * i + 1];
    DS_CHECK(r >= 0 && r < M && c >= 0 && c < M, "ds_plan_create_coo: index (%lld,%lld) outside [0,%lld)",
             (long long)r, (long long)c, (long long)M);
  }
  HostCsr A, At;
  coo_to_csr(M, nnz, indices, 2, indices + 1, 2, values, A);
  coo_to_csr(M, nnz, indices + 1, 2, indices, 2, values, At);
  // symmetric up to rounding: same pattern and |a_ij - a_ji| <= 1e-6 max|a| (a normalised Laplacian assembled in
  // floating point is symmetric only to an ulp; L~^T is then served by the same tables, a <= 1e-6 perturbation)
  bool sym = (A.rowptr == At.rowptr) && (A.col == At.col);
  if (sym && nnz > 0) {
    float amax = 0.f;
    for (int64_t i = 0; i < nnz; ++i) amax = std::max(amax, std::fabs(A.val[i]));
    for (int64_t i = 0; i < nnz && sym; ++i) sym = std::fabs(A.val[i] - At.val[i]) <= 1e-6f * amax;
  }

  ds_plan* P = new ds_plan();
  P->M = M;
  P->nnz = nnz;
  P->symmetric = sym;
  cudaGetDevice(&P->device);
  int rc = build_sparse_dev(A, ell_width, P->fwd, P->device_bytes);
  if (rc == 0) {
    if (sym) P->bwd = P->fwd;
    else rc = build_sparse_dev(At, ell_width, P->bwd, P->device_bytes);
  }
  if (rc != 0) {
    ds_plan_destroy(P);
    return rc;
  }
  *plan_out = P;
  return 0;
}

int ds_plan_destroy(ds_plan_t* plan) {
  if (!plan) return 0;
  if (plan->lattice) ds::lattice_free(plan->lattice);
  if (!plan->symmetric) ds::free_sparse_dev(plan->bwd);
  ds::free_sparse_dev(plan->fwd);
  delete plan;
  return 0;
}

int ds_plan_info(const ds_plan_t* plan, int32_t what, int64_t* value_out) {
  using namespace ds;
  DS_CHECK(plan && value_out, "ds_plan_info: NULL argument");
  switch (what) {
    case 0: *value_out = plan->M; break;
    case 1: *value_out = plan->nnz; break;
    case 2: *value_out = plan->fwd.W; break;
    case 3: *value_out = plan->fwd.n_tail_rows; break;
    case 4: *value_out = plan->fwd.tail_nnz; break;
    case 5: *value_out = plan->bwd.W; break;
    case 6: *value_out = plan->bwd.n_tail_rows; break;
    case 7: *value_out = plan->device_bytes; break;
    case 8: *value_out = plan->symmetric ? 1 : 0; break;
    case 9: *value_out = plan->lattice ? 1 : 0; break;
    default: return fail("ds_plan_info: unknown selector %d", what);
  }
  return 0;
}

}  // extern "C"
