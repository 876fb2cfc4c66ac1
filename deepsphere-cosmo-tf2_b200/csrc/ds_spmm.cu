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
#include "ds_ptx.cuh"

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

// Main path (F % 4 == 0, F <= 128): bulk-async staged hop.
// A CTA walks "units" of U consecutive rows of one batch element (a compact NESTED patch).  One producer
// thread streams each unit's contiguous slabs - its rows of `in`, of `prev` / `add`, and its packed ELL
// words - into shared memory with cp.async.bulk, NSTAGE units ahead, completing on an mbarrier; the
// consumer warps gather from shared memory and only go to global (L1/L2) for neighbours outside the unit.
// Bytes in flight are therefore set by the pipeline depth (>= 150 KB per SM), not by how many loads a
// thread can keep pending - which is what capped the register-gather kernels at ~3 TB/s (r1c bw_probe).
struct TileStage {
  uint64_t full;
  uint64_t empty;
};

__global__ void __launch_bounds__(544, 1) spmm_tile_kernel(const int4* __restrict__ ell_pk, int NP, int64_t M, int64_t B,
                                                           int FV, const float4* __restrict__ in, float alpha,
                                                           const float4* __restrict__ prev, float beta,
                                                           const float4* __restrict__ add, float gamma,
                                                           float4* __restrict__ out, int lpr_log2, int U, int nstage) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t in_bytes = (uint32_t)U * FV * 16;
  const uint32_t ell_bytes = (uint32_t)U * NP * 16;
  const uint32_t n_slabs = 1u + (prev ? 1u : 0u) + (add ? 1u : 0u);
  const uint32_t stage_bytes = n_slabs * in_bytes + ell_bytes;
  TileStage* bars = reinterpret_cast<TileStage*>(smem + (size_t)nstage * stage_bytes);
  const int n_consumers = blockDim.x - 32;
  const int64_t units_per_b = (M + U - 1) / U;
  const int64_t n_units = B * units_per_b;

  if (threadIdx.x == 0) {
    for (int s = 0; s < nstage; ++s) {
      ptx::mbar_init(&bars[s].full, 1);
      ptx::mbar_init(&bars[s].empty, n_consumers);
    }
    ptx::fence_mbar_init();
  }
  __syncthreads();

  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) {
      uint32_t it = 0;
      for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x, ++it) {
        const int s = it % nstage;
        ptx::mbar_wait(&bars[s].empty, ((it / nstage) & 1) ^ 1);
        const int64_t b = u / units_per_b;
        const int64_t m0 = (u - b * units_per_b) * U;
        const uint32_t rows = (uint32_t)min((int64_t)U, M - m0);
        const uint32_t nb = rows * FV * 16, eb = rows * NP * 16;
        uint8_t* st = smem + (size_t)s * stage_bytes;
        ptx::mbar_arrive_expect_tx(&bars[s].full, n_slabs * nb + eb);
        const int64_t goff = (b * M + m0) * FV;
        ptx::bulk_load_1d(st, in + goff, nb, &bars[s].full);
        uint32_t o = in_bytes;
        if (prev) { ptx::bulk_load_1d(st + o, prev + goff, nb, &bars[s].full); o += in_bytes; }
        if (add) { ptx::bulk_load_1d(st + o, add + goff, nb, &bars[s].full); o += in_bytes; }
        ptx::bulk_load_1d(st + o, ell_pk + m0 * NP, eb, &bars[s].full);
      }
    }
    return;
  }

  const int t = threadIdx.x - 32;
  const int groups = n_consumers >> lpr_log2;
  const int g = t >> lpr_log2;
  const int c = t & ((1 << lpr_log2) - 1);
  const bool active = c < FV;
  uint32_t it = 0;
  for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x, ++it) {
    const int s = it % nstage;
    ptx::mbar_wait(&bars[s].full, (it / nstage) & 1);
    const int64_t b = u / units_per_b;
    const int m0 = (int)((u - b * units_per_b) * U);
    const int rows = (int)min((int64_t)U, M - m0);
    const uint8_t* st = smem + (size_t)s * stage_bytes;
    const float4* s_in = reinterpret_cast<const float4*>(st);
    const float4* s_prev = reinterpret_cast<const float4*>(st + in_bytes);
    const float4* s_add = reinterpret_cast<const float4*>(st + (prev ? 2 : 1) * in_bytes);
    const int4* s_ell = reinterpret_cast<const int4*>(st + n_slabs * in_bytes);
    const float4* inb = in + b * M * FV;
    float4* outb = out + (b * M + m0) * FV;
    if (active) {
      for (int r = g; r < rows; r += groups) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int4* row = s_ell + r * NP;
#pragma unroll 2
        for (int n = 0; n < NP; ++n) {
          const int4 q = row[n];
          const uint32_t l0 = (uint32_t)(q.x - m0), l1 = (uint32_t)(q.z - m0);
          const float4 x0 = l0 < (uint32_t)rows ? s_in[l0 * FV + c] : __ldg(inb + (uint32_t)(q.x * FV + c));
          const float4 x1 = l1 < (uint32_t)rows ? s_in[l1 * FV + c] : __ldg(inb + (uint32_t)(q.z * FV + c));
          const float w0 = __int_as_float(q.y), w1 = __int_as_float(q.w);
          acc.x = fmaf(w0, x0.x, acc.x); acc.y = fmaf(w0, x0.y, acc.y);
          acc.z = fmaf(w0, x0.z, acc.z); acc.w = fmaf(w0, x0.w, acc.w);
          acc.x = fmaf(w1, x1.x, acc.x); acc.y = fmaf(w1, x1.y, acc.y);
          acc.z = fmaf(w1, x1.z, acc.z); acc.w = fmaf(w1, x1.w, acc.w);
        }
        acc.x *= alpha; acc.y *= alpha; acc.z *= alpha; acc.w *= alpha;
        if (prev != nullptr) {
          const float4 pv = s_prev[r * FV + c];
          acc.x = fmaf(beta, pv.x, acc.x); acc.y = fmaf(beta, pv.y, acc.y);
          acc.z = fmaf(beta, pv.z, acc.z); acc.w = fmaf(beta, pv.w, acc.w);
        }
        if (add != nullptr) {
          const float4 av = s_add[r * FV + c];
          acc.x = fmaf(gamma, av.x, acc.x); acc.y = fmaf(gamma, av.y, acc.y);
          acc.z = fmaf(gamma, av.z, acc.z); acc.w = fmaf(gamma, av.w, acc.w);
        }
        __stcs(outb + r * FV + c, acc);
      }
    }
    ptx::mbar_arrive(&bars[s].empty);  // this thread is done reading the stage
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

// Narrow rows (F <= 8 and not a multiple of 4: the 1 - 5 channel layers of the reference's notebooks): ONE thread per
// (sample, row) accumulates all F channels, so the packed (column, value) pairs of the row are read once instead of once
// per channel lane and no lane idles (spmm_ell_kernel<1> with F = 5 keeps 5 of 8 lanes busy and issues 3 loads per FMA:
// 60 us per hop at nside 64, batch 16 - 43 % of the quick_start training step).  Consecutive threads own consecutive rows.
template <int FMAX>
__global__ void __launch_bounds__(256) spmm_rowthread_kernel(const int4* __restrict__ ell_pk, int WH, int64_t M, int64_t B,
                                                             int F, const float* __restrict__ in, float alpha,
                                                             const float* __restrict__ prev, float beta,
                                                             const float* __restrict__ add, float gamma,
                                                             float* __restrict__ out) {
  const int64_t total = B * M;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / M, m = idx - b * M;
    const float* inb = in + b * M * F;
    const int4* row = ell_pk + m * WH;
    float acc[FMAX];
#pragma unroll
    for (int f = 0; f < FMAX; ++f) acc[f] = 0.f;
    for (int n = 0; n < WH; ++n) {
      const int4 e = __ldg(row + n);
      const float* x0 = inb + (int64_t)e.x * F;
      const float* x1 = inb + (int64_t)e.z * F;
      const float w0 = __int_as_float(e.y), w1 = __int_as_float(e.w);
#pragma unroll
      for (int f = 0; f < FMAX; ++f)
        if (f < F) acc[f] = fmaf(w0, __ldg(x0 + f), acc[f]);
#pragma unroll
      for (int f = 0; f < FMAX; ++f)
        if (f < F) acc[f] = fmaf(w1, __ldg(x1 + f), acc[f]);
    }
    const int64_t off = idx * F;
#pragma unroll
    for (int f = 0; f < FMAX; ++f)
      if (f < F) {
        float v = acc[f] * alpha;
        if (prev != nullptr) v = fmaf(beta, __ldg(prev + off + f), v);
        if (add != nullptr) v = fmaf(gamma, __ldg(add + off + f), v);
        out[off + f] = v;
      }
  }
}

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
  // small problems (the closure sub-problem of the irregular lattice tiles, the coarse levels of a network: the whole
  // hop lives in L2) take the plain gather kernel: the bulk-copy ring of spmm_tile_kernel costs ~19 us per launch in
  // pipeline fill alone (32 such launches were 10 % of a HealpyGCNN training step, profiles/r2p_launches_model_train.csv)
  static const int64_t small_bytes = [] { const char* e = getenv("DEEPSPHERE_SPMM_SMALL_BYTES"); return e ? atoll(e) : (int64_t)(24 << 20); }();
  const bool small = B * S.M * F * 4 <= small_bytes;
  if (vec4 && F <= 128 && S.M * (F / 4) < (int64_t)1 << 31 && S.Wp <= 64 && !small) {
    // unit: rows so that one slab is <= 32 KB; stages from the shared-memory budget
    int U = 1024;
    while (U > 16 && (int64_t)U * F * 4 > 32768) U >>= 1;
    const int n_slabs = 1 + (prev ? 1 : 0) + (add ? 1 : 0);
    const size_t stage_bytes = (size_t)n_slabs * U * F * 4 + (size_t)U * (S.Wp / 2) * 16;
    int dev = 0, max_smem = 0;
    DS_CUDA(cudaGetDevice(&dev));
    DS_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    int nstage = (int)std::min<size_t>(4, (max_smem - 1024) / stage_bytes);
    if (nstage >= 2) {
      static PerDeviceOnce attr_once;
      DS_TRY(attr_once.run([&]() -> int {
        DS_CUDA(cudaFuncSetAttribute(spmm_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        return 0;
      }));
      const int64_t upb = (S.M + U - 1) / U;
      const int64_t nblk = std::max<int64_t>(1, std::min<int64_t>(B * upb, num_sms()));
      spmm_tile_kernel<<<(unsigned)nblk, 544, nstage * stage_bytes + nstage * sizeof(TileStage) + 64, st>>>(
          S.ell_pk, S.Wp / 2, S.M, B, (int)(F / 4), reinterpret_cast<const float4*>(in), alpha,
          reinterpret_cast<const float4*>(prev), prev ? beta : 0.f, reinterpret_cast<const float4*>(add),
          add ? gamma : 0.f, reinterpret_cast<float4*>(out), lpr_log2, U, nstage);
    } else {
      spmm_ell_pk_kernel<<<(unsigned)blocks, threads, 0, st>>>(
          S.ell_pk, S.Wp / 2, S.M, B, (int)(F / 4), reinterpret_cast<const float4*>(in), alpha,
          reinterpret_cast<const float4*>(prev), prev ? beta : 0.f, reinterpret_cast<const float4*>(add),
          add ? gamma : 0.f, reinterpret_cast<float4*>(out), lpr_log2, unit_rows);
    }
  } else if (vec4 && F <= 128 && S.M * (F / 4) < (int64_t)1 << 31) {
    spmm_ell_pk_kernel<<<(unsigned)blocks, threads, 0, st>>>(
        S.ell_pk, S.Wp / 2, S.M, B, (int)(F / 4), reinterpret_cast<const float4*>(in), alpha,
        reinterpret_cast<const float4*>(prev), prev ? beta : 0.f, reinterpret_cast<const float4*>(add),
        add ? gamma : 0.f, reinterpret_cast<float4*>(out), lpr_log2, unit_rows);
  } else if (!vec4 && F <= 8 && S.ell_pk != nullptr && S.Wp >= 2) {
    const int64_t nb = std::max<int64_t>(1, std::min<int64_t>((B * S.M + 255) / 256, max_blocks));
    spmm_rowthread_kernel<8><<<(unsigned)nb, 256, 0, st>>>(S.ell_pk, S.Wp / 2, S.M, B, (int)F, in, alpha, prev,
                                                          prev ? beta : 0.f, add, add ? gamma : 0.f, out);
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
