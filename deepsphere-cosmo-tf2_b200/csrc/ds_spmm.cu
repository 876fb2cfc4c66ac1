// One hop of the graph recursion as a streaming ELL SpMM (+ CSR tail):
//   out[b,m,:] = alpha * sum_n val[m,n] * in[b, col[m,n], :] + beta * prev[b,m,:] + gamma * add[b,m,:]
// This is the generic (any L, any F) path that replaces tf.sparse.sparse_dense_matmul
// (utils.py:76) and the `2*... - x0` elementwise pass (gnn_layers.py:141) in one kernel,
// in the reference's native [B, M, F] layout (no transposes, gnn_layers.py:131-132).
//
// Mapping: a group of LPR lanes (power of two <= 32) owns one (b, m) row and strides over
// the row's feature vector in float4 (F % 4 == 0) or scalar steps, so global loads of a
// gathered row are contiguous.  Rows are walked in (b, m) order: consecutive groups gather
// from neighbouring pixels (NESTED order keeps graph neighbours close in memory), so the
// 9x re-use of each input row is served by L1/L2.  HBM-bound: algorithmic traffic is
// read in + read prev + write out = 3 * B*M*F*4 bytes per hop (plan: M*W*8 bytes).
#include "ds_common.cuh"

namespace ds {
namespace {

template <int V>
struct Vec;
template <>
struct Vec<1> {
  float v[1];
  __device__ __forceinline__ static Vec load(const float* p) { Vec r; r.v[0] = __ldg(p); return r; }
  __device__ __forceinline__ void store(float* p) const { p[0] = v[0]; }
};
template <>
struct Vec<4> {
  float v[4];
  __device__ __forceinline__ static Vec load(const float* p) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    Vec r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

// streaming (evict-first) access for operands that are touched exactly once per hop
template <int V>
__device__ __forceinline__ Vec<V> load_stream(const float* p);
template <>
__device__ __forceinline__ Vec<1> load_stream<1>(const float* p) { Vec<1> r; r.v[0] = __ldcs(p); return r; }
template <>
__device__ __forceinline__ Vec<4> load_stream<4>(const float* p) {
  const float4 t = __ldcs(reinterpret_cast<const float4*>(p));
  Vec<4> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
}
template <int V>
__device__ __forceinline__ void store_stream(float* p, const Vec<V>& a);
template <>
__device__ __forceinline__ void store_stream<1>(float* p, const Vec<1>& a) { __stcs(p, a.v[0]); }
template <>
__device__ __forceinline__ void store_stream<4>(float* p, const Vec<4>& a) {
  __stcs(reinterpret_cast<float4*>(p), make_float4(a.v[0], a.v[1], a.v[2], a.v[3]));
}

// A CTA owns `unit_rows` CONSECUTIVE rows of one batch element at a time.  In NESTED order that is a
// compact patch of the sphere, so the ~9 gathers per row hit rows the same CTA (same SM) has just
// touched: the re-use is served by L1 instead of L2, and HBM sees each input row about once.
template <int V>
__global__ void __launch_bounds__(256, 3) spmm_ell_kernel(const int32_t* __restrict__ ell_col,
                                                          const float* __restrict__ ell_val, int W, int64_t M,
                                                          int64_t B, int64_t F, const float* __restrict__ in,
                                                          float alpha, const float* __restrict__ prev, float beta,
                                                          const float* __restrict__ add, float gamma,
                                                          float* __restrict__ out, int lpr_log2, int unit_rows) {
  constexpr int CH = 9;  // gathers kept in flight per thread (= the HEALPix row length: 8 neighbours + diagonal)
  const int lpr = 1 << lpr_log2;
  const int groups_per_block = blockDim.x >> lpr_log2;
  const int g = threadIdx.x >> lpr_log2;
  const int sub = threadIdx.x & (lpr - 1);
  const int64_t FV = F / V;
  const int64_t units_per_b = (M + unit_rows - 1) / unit_rows;
  const int64_t n_units = B * units_per_b;
  for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
    const int64_t b = u / units_per_b;
    const int64_t m0 = (u - b * units_per_b) * unit_rows;
    const int64_t m1 = min(M, m0 + (int64_t)unit_rows);
    const float* inb = in + b * M * F;
    for (int64_t m = m0 + g; m < m1; m += groups_per_block) {
      const int32_t* cols = ell_col + m * W;
      const float* vals = ell_val + m * W;
      for (int64_t c = sub; c < FV; c += lpr) {
        const int64_t off = (b * M + m) * F + c * V;
        // the two streamed operands first, so their latency overlaps the gathers
        Vec<V> pv, av;
        if (prev != nullptr) pv = load_stream<V>(prev + off);
        if (add != nullptr) av = load_stream<V>(add + off);
        Vec<V> acc;
#pragma unroll
        for (int i = 0; i < V; ++i) acc.v[i] = 0.f;
        for (int n0 = 0; n0 < W; n0 += CH) {
          int32_t col[CH];
          float w[CH];
#pragma unroll
          for (int i = 0; i < CH; ++i) {
            const bool ok = n0 + i < W;
            col[i] = ok ? __ldg(cols + n0 + i) : (int32_t)m;
            w[i] = ok ? __ldg(vals + n0 + i) : 0.f;
          }
          Vec<V> xv[CH];
#pragma unroll
          for (int i = 0; i < CH; ++i) xv[i] = Vec<V>::load(inb + (int64_t)col[i] * F + c * V);
#pragma unroll
          for (int i = 0; i < CH; ++i)
#pragma unroll
            for (int e = 0; e < V; ++e) acc.v[e] = fmaf(w[i], xv[i].v[e], acc.v[e]);
        }
#pragma unroll
        for (int i = 0; i < V; ++i) acc.v[i] *= alpha;
        if (prev != nullptr) {
#pragma unroll
          for (int i = 0; i < V; ++i) acc.v[i] = fmaf(beta, pv.v[i], acc.v[i]);
        }
        if (add != nullptr) {
#pragma unroll
          for (int i = 0; i < V; ++i) acc.v[i] = fmaf(gamma, av.v[i], acc.v[i]);
        }
        store_stream<V>(out + off, acc);
      }
    }
  }
}

// Fast path (F % 4 == 0, F <= 128): packed ELL - one 16-byte load brings two (column, value) entries -
// 32-bit index arithmetic and all gathers of a 10-entry chunk in flight before the FMAs.  Roughly half
// the instructions per row of the generic kernel above, which ncu showed to be issue-bound (r1b).
__global__ void __launch_bounds__(256, 2) spmm_ell_pk_kernel(const int4* __restrict__ ell_pk, int NP, int64_t M,
                                                             int64_t B, int FV, const float4* __restrict__ in,
                                                             float alpha, const float4* __restrict__ prev, float beta,
                                                             const float4* __restrict__ add, float gamma,
                                                             float4* __restrict__ out, int lpr_log2, int unit_rows) {
  constexpr int CHP = 5;  // 16-byte words per chunk = 10 entries (HEALPix rows: 9 + 1 pad)
  const int groups_per_block = blockDim.x >> lpr_log2;
  const int g = threadIdx.x >> lpr_log2;
  const int c = threadIdx.x & ((1 << lpr_log2) - 1);
  const int64_t units_per_b = (M + unit_rows - 1) / unit_rows;
  const int64_t n_units = B * units_per_b;
  if (c >= FV) return;
  for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
    const int64_t b = u / units_per_b;
    const int m0 = (int)((u - b * units_per_b) * unit_rows);
    const int m1 = (int)min(M, (int64_t)m0 + unit_rows);
    const float4* inb = in + b * M * FV;
    for (int m = m0 + g; m < m1; m += groups_per_block) {
      const int64_t off = (b * M + m) * FV + c;
      float4 pv, av;
      if (prev != nullptr) pv = __ldcs(prev + off);
      if (add != nullptr) av = __ldcs(add + off);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      const int4* row = ell_pk + (int64_t)m * NP;
      for (int n0 = 0; n0 < NP; n0 += CHP) {
        int4 q[CHP];
#pragma unroll
        for (int i = 0; i < CHP; ++i) q[i] = n0 + i < NP ? __ldg(row + n0 + i) : make_int4(m, 0, m, 0);
        float4 x0[CHP], x1[CHP];
#pragma unroll
        for (int i = 0; i < CHP; ++i) {
          x0[i] = __ldg(inb + (uint32_t)(q[i].x * FV + c));
          x1[i] = __ldg(inb + (uint32_t)(q[i].z * FV + c));
        }
#pragma unroll
        for (int i = 0; i < CHP; ++i) {
          const float w0 = __int_as_float(q[i].y), w1 = __int_as_float(q[i].w);
          acc.x = fmaf(w0, x0[i].x, acc.x); acc.y = fmaf(w0, x0[i].y, acc.y);
          acc.z = fmaf(w0, x0[i].z, acc.z); acc.w = fmaf(w0, x0[i].w, acc.w);
          acc.x = fmaf(w1, x1[i].x, acc.x); acc.y = fmaf(w1, x1[i].y, acc.y);
          acc.z = fmaf(w1, x1[i].z, acc.z); acc.w = fmaf(w1, x1[i].w, acc.w);
        }
      }
      acc.x *= alpha; acc.y *= alpha; acc.z *= alpha; acc.w *= alpha;
      if (prev != nullptr) {
        acc.x = fmaf(beta, pv.x, acc.x); acc.y = fmaf(beta, pv.y, acc.y);
        acc.z = fmaf(beta, pv.z, acc.z); acc.w = fmaf(beta, pv.w, acc.w);
      }
      if (add != nullptr) {
        acc.x = fmaf(gamma, av.x, acc.x); acc.y = fmaf(gamma, av.y, acc.y);
        acc.z = fmaf(gamma, av.z, acc.z); acc.w = fmaf(gamma, av.w, acc.w);
      }
      __stcs(out + off, acc);
    }
  }
}

// rows longer than the ELL width: out[b,row,:] += alpha * sum_tail val * in[b,col,:]
__global__ void __launch_bounds__(256) spmm_tail_kernel(const int32_t* __restrict__ tail_rows,
                                                        const int64_t* __restrict__ tail_rowptr,
                                                        const int32_t* __restrict__ tail_col,
                                                        const float* __restrict__ tail_val, int64_t n_tail, int64_t M,
                                                        int64_t B, int64_t F, const float* __restrict__ in, float alpha,
                                                        float* __restrict__ out, int lpr_log2) {
  const int lpr = 1 << lpr_log2;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n_groups = ((int64_t)gridDim.x * blockDim.x) >> lpr_log2;
  const int sub = (int)(tid & (lpr - 1));
  const int64_t R = B * n_tail;
  for (int64_t r = tid >> lpr_log2; r < R; r += n_groups) {
    const int64_t b = r / n_tail;
    const int64_t t = r - b * n_tail;
    const int64_t m = tail_rows[t];
    const int64_t p0 = tail_rowptr[t], p1 = tail_rowptr[t + 1];
    const float* inb = in + b * M * F;
    for (int64_t c = sub; c < F; c += lpr) {
      float acc = 0.f;
      for (int64_t p = p0; p < p1; ++p) acc = fmaf(__ldg(tail_val + p), __ldg(inb + (int64_t)tail_col[p] * F + c), acc);
      out[(b * M + m) * F + c] += alpha * acc;
    }
  }
}

inline int ilog2_ceil(int64_t v) {
  int l = 0;
  while ((1LL << l) < v) ++l;
  return l;
}

}  // namespace

int launch_spmm(const SparseDev& S, int64_t B, int64_t F, const float* in, float alpha, const float* prev, float beta,
                const float* add, float gamma, float* out, cudaStream_t st) {
  DS_CHECK(B > 0 && F > 0, "spmm: empty batch or feature dimension");
  DS_CHECK(in != out, "spmm: in-place hop is not supported (gather hazard)");
  auto aligned = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec4 = (F % 4 == 0) && aligned(in) && aligned(prev) && aligned(add) && aligned(out);
  const int V = vec4 ? 4 : 1;
  const int lpr_log2 = std::min(5, ilog2_ceil(F / V));
  // unit = rows per CTA visit: ~64 KB of input rows (a square NESTED patch), so two resident CTAs'
  // patches + halos fit L1
  int unit_rows = 64;
  while (unit_rows < 4096 && (int64_t)unit_rows * 4 * F * 4 <= 65536) unit_rows *= 4;
  const int64_t units_per_b = (S.M + unit_rows - 1) / unit_rows;
  const int64_t max_blocks = (int64_t)num_sms() * 3 * 8;
  const int64_t blocks = std::max<int64_t>(1, std::min<int64_t>(B * units_per_b, max_blocks));
  const int threads = 256;
  if (vec4 && F <= 128 && S.M * (F / 4) < (int64_t)1 << 31) {
    spmm_ell_pk_kernel<<<(unsigned)blocks, threads, 0, st>>>(
        S.ell_pk, S.Wp / 2, S.M, B, (int)(F / 4), reinterpret_cast<const float4*>(in), alpha,
        reinterpret_cast<const float4*>(prev), prev ? beta : 0.f, reinterpret_cast<const float4*>(add),
        add ? gamma : 0.f, reinterpret_cast<float4*>(out), lpr_log2, unit_rows);
  } else if (vec4) {
    spmm_ell_kernel<4><<<(unsigned)blocks, threads, 0, st>>>(S.ell_col, S.ell_val, S.W, S.M, B, F, in, alpha, prev,
                                                             prev ? beta : 0.f, add, add ? gamma : 0.f, out, lpr_log2,
                                                             unit_rows);
  } else {
    spmm_ell_kernel<1><<<(unsigned)blocks, threads, 0, st>>>(S.ell_col, S.ell_val, S.W, S.M, B, F, in, alpha, prev,
                                                             prev ? beta : 0.f, add, add ? gamma : 0.f, out, lpr_log2,
                                                             unit_rows);
  }
  DS_LAUNCHED();
  if (S.n_tail_rows > 0) {
    const int tl = std::min(5, ilog2_ceil(F));
    const int64_t gpb = 256 >> tl;
    int64_t tb = (B * S.n_tail_rows + gpb - 1) / gpb;
    if (tb > max_blocks) tb = max_blocks;
    spmm_tail_kernel<<<(unsigned)tb, 256, 0, st>>>(S.tail_rows, S.tail_rowptr, S.tail_col, S.tail_val,
                                                       S.n_tail_rows, S.M, B, F, in, alpha, out, tl);
    DS_LAUNCHED();
  }
  return 0;
}

}  // namespace ds

extern "C" int ds_spmm(const ds_plan_t* plan, int32_t transpose, int64_t B, int64_t F, const float* in, float alpha,
                       const float* prev, float beta, const float* add, float gamma, float* out, void* stream) {
  using namespace ds;
  DS_CHECK(plan && in && out, "ds_spmm: NULL argument");
  return launch_spmm(transpose ? plan->bwd : plan->fwd, B, F, in, alpha, prev, beta, add, gamma, out,
                     (cudaStream_t)stream);
}
