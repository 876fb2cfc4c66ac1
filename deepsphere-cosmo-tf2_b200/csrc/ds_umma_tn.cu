// tcgen05 weight-gradient kernel for DS_MODE_TF32 / DS_MODE_TF32X3:
//   dW[(kc*c_kc_stride + seg*c_seg_stride), n] = sum_r A_seg[r, kc] * D[r, n]        (reduction over R rows)
// i.e. dkernel = X_stack^T * dY of the contraction at gnn_layers.py:149, with the K basis
// tensors read in place.  The reduction dimension r is the *strided* one for both operands, so both
// are fed MN-major.  For 32-bit (tf32) MN-major operands the only UMMA shared-memory layout is the
// "128-byte swizzle with 32-byte atoms" (UMMA layout type SWIZZLE_128B_BASE32B: 4 k-rows x 128 B atoms,
// 32-byte chunk index XOR k-row % 4), which is exactly what a TMA box {32 channels, 16 rows} written with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B produces - no transposes anywhere.
//
// Work split: each persistent CTA reduces a contiguous range of rows for ALL nseg*Kc x N outputs,
// accumulated in TMEM (ceil(nseg*ceil(Kc/32)/4) accumulators of 128 lanes x N columns), so every input
// row is read from HBM exactly once; per-CTA partials are then summed by a small second kernel
// (deterministic, no atomics).  HBM-bound: (nseg*Kc + N)*4 bytes per row.
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "ds_common.cuh"
#include "ds_ptx.cuh"

namespace ds {
namespace {

constexpr int BKR = 16;          // rows (reduction steps) per stage = two K=8 UMMA steps
constexpr int BLK = 32;          // channels per MN block = one 128-byte swizzle row
constexpr int BLK_BYTES = BKR * BLK * 4;  // 2 KB: one MN block of one stage
constexpr int MAX_STAGES = 8;

struct TnParams {
  int64_t R;
  int N;            // columns of D (UMMA N)
  int nb_blocks;    // ceil(N / 32)
  int n_q;          // MN blocks of the A side = nseg * ceil(Kc / 32)
  int q_per_seg;    // ceil(Kc / 32)
  int m_tiles;      // ceil(n_q / 4)
  int three_pass;
  int stages;
  uint32_t stage_bytes;
  uint32_t tmem_cols;
  int64_t rblocks_per_cta;
  float* partial;   // [gridDim.x][n_q * 32][N]
  // raw operand pointers for the cp.async producers
  const float* A0;     // segment 0, [R, Kc]
  const float* Arest;  // segments 1.., [nseg-1, R, Kc]
  const float* D;      // [R, N]
  int64_t Kc;
  int use_cp_async;
};

struct TnCtl {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t split_done[MAX_STAGES];
  uint64_t acc_full;
  uint32_t tmem_base;
};

// ---- cp.async operand producers of umma_gemm_tn_kernel (warps 2-5) ---------------------------------------------------
// The TMA box in 32-byte-atom swizzle mode sustains only ~10 B/clk/SM (ncu: 2.6 TB/s chip-wide, nothing saturated); 128
// threads issuing 16-byte cp.async straight into the same layout (row r of a 16 x 32 block at r*128 B, 32-byte chunk index
// XOR r % 4) reach the HBM rate.  One 16-byte chunk per thread per block; out-of-range rows / channels are zero-filled.
//
// One producer warp walks EVERY stage, so its instruction chain per stage is the pace of the kernel: with the block count
// a run-time value the 16 block slots were predicated, not skipped - 317 issue slots per 16-row stage (three integer
// divisions for the stage / phase indices, 64-bit selects and size arithmetic per slot), ~1 300 cycles per stage whatever
// the stage carried: 5.9 TB/s with 64-channel operands but 1.8 TB/s with 16 (profiles/r2ac_umma_gemm_tn_narrow_summary.txt).
// NB is now a template parameter (exactly NB copies, no predicates), everything that does not change from stage to stage
// lives in registers (the running source address of this thread's chunk in every block, its advance per stage, whether
// it is zero-filled), the stage / phase indices are counters, and the zero-fill uses cp.async's ignore-src predicate.
template <int NB>
__device__ __forceinline__ void tn_produce(const TnParams& p, TnCtl* ctl, uint8_t* stage_base, int64_t rb0, int64_t n_it) {
  const int t = threadIdx.x - 64;
  const int r = t >> 3, j = t & 7;
  const uint32_t dst_off = (uint32_t)(r * 128 + (((j >> 1) ^ (r & 3)) * 32) + (j & 1) * 16);
  constexpr int LAG = 3;  // stages whose copies may still be in flight per thread
  const char* src[NB];
  uint32_t step[NB], ign[NB];
  const int64_t row0 = rb0 * BKR + r;  // this thread's row in the first stage
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    src[k] = reinterpret_cast<const char*>(p.A0);
    step[k] = 0;
    ign[k] = 1;
    if (k < p.n_q) {
      const int seg = k / p.q_per_seg, cb = k % p.q_per_seg;
      const int col = cb * BLK + 4 * j;
      const float* base = seg == 0 ? p.A0 : p.Arest + (int64_t)(seg - 1) * p.R * p.Kc;
      if (col < p.Kc && row0 < p.R) {
        src[k] = reinterpret_cast<const char*>(base + row0 * p.Kc + col);
        step[k] = (uint32_t)(BKR * p.Kc * 4);
        ign[k] = 0;
      }
    } else {
      const int col = (k - p.n_q) * BLK + 4 * j;
      if (col < p.N && row0 < p.R) {
        src[k] = reinterpret_cast<const char*>(p.D + row0 * (int64_t)p.N + col);
        step[k] = (uint32_t)(BKR * p.N * 4);
        ign[k] = 0;
      }
    }
  }
  // stages in which this thread's row exists (only the last stage of the whole problem can be ragged)
  const int64_t live = row0 < p.R ? (p.R - row0 + BKR - 1) / BKR : 0;
  const uint32_t stage0 = ptx::smem_u32(stage_base) + dst_off;
  int s = 0, s_done = 0;  // stage of iteration `it` / of iteration it - LAG
  uint32_t ph = 0, st = stage0;
  for (int64_t it = 0; it < n_it + LAG; ++it) {
    if (it < n_it) {
      if (it == live) {  // the row ran past the end: zero-fill from a valid address from here on
#pragma unroll
        for (int k = 0; k < NB; ++k) {
          src[k] = reinterpret_cast<const char*>(p.A0);
          step[k] = 0;
          ign[k] = 1;
        }
      }
      ptx::mbar_wait(&ctl->empty[s], ph ^ 1);
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        // the D blocks follow the A blocks in the stage (off_d = n_q * BLK_BYTES)
        asm volatile(
            "{\n\t.reg .pred pz;\n\tsetp.ne.u32 pz, %2, 0;\n\t"
            "cp.async.cg.shared.global [%0], [%1], 16, pz;\n\t}" ::"r"(st + (uint32_t)k * BLK_BYTES),
            "l"(src[k]), "r"(ign[k])
            : "memory");
        src[k] += step[k];
      }
      st += p.stage_bytes;
      if (++s == p.stages) { s = 0; ph ^= 1; st = stage0; }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (it >= LAG) {
      asm volatile("cp.async.wait_group %0;" ::"n"(LAG) : "memory");
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&ctl->full[s_done]);
      if (++s_done == p.stages) s_done = 0;
    }
  }
}

__global__ void __launch_bounds__(320, 1)
umma_gemm_tn_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_arest,
                    const __grid_constant__ CUtensorMap map_d, const TnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* stage_base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  TnCtl* ctl = reinterpret_cast<TnCtl*>(stage_base + (size_t)p.stages * p.stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_rblocks = (p.R + BKR - 1) / BKR;
  const int64_t rb0 = (int64_t)blockIdx.x * p.rblocks_per_cta;
  const int64_t rb1 = min(n_rblocks, rb0 + p.rblocks_per_cta);
  const int64_t n_it = rb1 > rb0 ? rb1 - rb0 : 0;
  // stage layout: [A hi: n_q blocks][D hi: nb blocks][A lo][D lo]  (each block 2 KB)
  const uint32_t blocks_hi = (uint32_t)(p.n_q + p.nb_blocks);
  const uint32_t off_d = (uint32_t)p.n_q * BLK_BYTES;
  const uint32_t off_lo = blocks_hi * BLK_BYTES;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&ctl->full[s], p.use_cp_async ? 128 : 1);
      ptx::mbar_init(&ctl->empty[s], 1);
      ptx::mbar_init(&ctl->split_done[s], 128);
    }
    ptx::mbar_init(&ctl->acc_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tma_prefetch_desc(&map_a0);
    ptx::tma_prefetch_desc(&map_arest);
    ptx::tma_prefetch_desc(&map_d);
  }
  if (warp == 1) ptx::tmem_alloc(&ctl->tmem_base, p.tmem_cols);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    if (!p.use_cp_async && ptx::elect_one()) {
      for (int64_t it = 0; it < n_it; ++it) {
        const int s = (int)(it % p.stages);
        const uint32_t ph = (uint32_t)(it / p.stages) & 1;
        ptx::mbar_wait(&ctl->empty[s], ph ^ 1);
        uint8_t* st = stage_base + (size_t)s * p.stage_bytes;
        ptx::mbar_arrive_expect_tx(&ctl->full[s], blocks_hi * BLK_BYTES);
        const int32_t r0 = (int32_t)((rb0 + it) * BKR);
        for (int q = 0; q < p.n_q; ++q) {
          const int seg = q / p.q_per_seg, cb = q % p.q_per_seg;
          if (seg == 0) ptx::tma_load_3d(st + q * BLK_BYTES, &map_a0, cb * BLK, r0, 0, &ctl->full[s]);
          else ptx::tma_load_3d(st + q * BLK_BYTES, &map_arest, cb * BLK, r0, seg - 1, &ctl->full[s]);
        }
        for (int nb = 0; nb < p.nb_blocks; ++nb)
          ptx::tma_load_3d(st + off_d + nb * BLK_BYTES, &map_d, nb * BLK, r0, 0, &ctl->full[s]);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = ptx::make_idesc_tf32(128, p.N, 1, 1);  // both operands MN-major
    for (int64_t it = 0; it < n_it; ++it) {
      const int s = (int)(it % p.stages);
      const uint32_t ph = (uint32_t)(it / p.stages) & 1;
      ptx::mbar_wait(&ctl->full[s], ph);
      if (p.three_pass) ptx::mbar_wait(&ctl->split_done[s], ph);
      ptx::tc_fence_after_sync();
      if (ptx::elect_one()) {
        const uint32_t st = ptx::smem_u32(stage_base + (size_t)s * p.stage_bytes);
        // MN-major, 32-byte-atom 128B swizzle: LBO = distance between 32-channel blocks, SBO = 512 (4 k-rows)
        const uint64_t d_hi = ptx::make_smem_desc(st + off_d, BLK_BYTES, 512, ptx::LAYOUT_SWIZZLE_128B_BASE32B);
        const uint64_t d_lo = ptx::make_smem_desc(st + off_lo + off_d, BLK_BYTES, 512, ptx::LAYOUT_SWIZZLE_128B_BASE32B);
        for (int mt = 0; mt < p.m_tiles; ++mt) {
          const uint64_t a_hi = ptx::make_smem_desc(st + mt * 4 * BLK_BYTES, BLK_BYTES, 512, ptx::LAYOUT_SWIZZLE_128B_BASE32B);
          const uint64_t a_lo =
              ptx::make_smem_desc(st + off_lo + mt * 4 * BLK_BYTES, BLK_BYTES, 512, ptx::LAYOUT_SWIZZLE_128B_BASE32B);
          const uint32_t d_tmem = tmem_base + (uint32_t)(mt * p.N);
#pragma unroll
          for (int j = 0; j < BKR / 8; ++j) {
            const uint64_t adv = (uint64_t)((j * 1024) >> 4);  // next 8 reduction rows
            const uint32_t first = (it == 0 && j == 0) ? 0u : 1u;
            if (p.three_pass) {
              ptx::umma_tf32(d_tmem, a_lo + adv, d_hi + adv, idesc, first);
              ptx::umma_tf32(d_tmem, a_hi + adv, d_lo + adv, idesc, 1u);
              ptx::umma_tf32(d_tmem, a_hi + adv, d_hi + adv, idesc, 1u);
            } else {
              ptx::umma_tf32(d_tmem, a_hi + adv, d_hi + adv, idesc, first);
            }
          }
        }
        ptx::umma_commit(&ctl->empty[s]);
        if (it == n_it - 1) ptx::umma_commit(&ctl->acc_full);
      }
      __syncwarp();
    }
  } else if (warp < 6) {
    if (p.use_cp_async) {
      // operand producers (the epilogue warps are idle during the main loop): tn_produce above
      switch (p.n_q + p.nb_blocks) {
#define DS_TN_CASE(nb) case nb: tn_produce<nb>(p, ctl, stage_base, rb0, n_it); break;
        DS_TN_CASE(2) DS_TN_CASE(3) DS_TN_CASE(4) DS_TN_CASE(5) DS_TN_CASE(6) DS_TN_CASE(7) DS_TN_CASE(8) DS_TN_CASE(9)
        DS_TN_CASE(10) DS_TN_CASE(11) DS_TN_CASE(12) DS_TN_CASE(13) DS_TN_CASE(14) DS_TN_CASE(15) DS_TN_CASE(16)
#undef DS_TN_CASE
        default: __trap();  // the launcher only takes this path for 2..16 blocks
      }
    }
    // epilogue: once, after the whole row range has been reduced
    const int qd = warp & 3;
    float* part = p.partial + (size_t)blockIdx.x * p.n_q * BLK * p.N;
    if (n_it > 0) {
      ptx::mbar_wait(&ctl->acc_full, 0);
      ptx::tc_fence_after_sync();
    }
    for (int mt = 0; mt < p.m_tiles; ++mt) {
      const int row = mt * 128 + qd * 32 + lane;  // output row = q*32 + channel
      for (int c0 = 0; c0 < p.N; c0 += 16) {
        uint32_t r[16];
        if (n_it > 0) {
          ptx::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(mt * p.N + c0), r);
          ptx::tmem_ld_wait();
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) r[e] = 0u;
        }
        if (row < p.n_q * BLK) {
#pragma unroll
          for (int v = 0; v < 4; ++v)
            *reinterpret_cast<float4*>(part + (size_t)row * p.N + c0 + v * 4) =
                make_float4(__uint_as_float(r[v * 4]), __uint_as_float(r[v * 4 + 1]), __uint_as_float(r[v * 4 + 2]),
                            __uint_as_float(r[v * 4 + 3]));
        }
      }
    }
  } else if (p.three_pass) {
    const int t = threadIdx.x - 6 * 32;
    const int n_vec = (int)(blocks_hi * BLK_BYTES / 16);
    for (int64_t it = 0; it < n_it; ++it) {
      const int s = (int)(it % p.stages);
      const uint32_t ph = (uint32_t)(it / p.stages) & 1;
      ptx::mbar_wait(&ctl->full[s], ph);
      float4* hi = reinterpret_cast<float4*>(stage_base + (size_t)s * p.stage_bytes);
      float4* lo = reinterpret_cast<float4*>(stage_base + (size_t)s * p.stage_bytes + off_lo);
      for (int idx = t; idx < n_vec; idx += 128) {
        const float4 v = hi[idx];
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
        h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
        h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
        h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
        // (high part: implicit hardware truncation, see ds_umma.cu)
        lo[idx] = l;
      }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&ctl->split_done[s]);
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// C[kc*s_kc + seg*s_seg + n*s_n] = sum over CTAs of partial[cta][q*32 + i][n],  kc = (q % q_per_seg)*32 + i
// 256 threads = 32 consecutive outputs x 8 slices of the CTA range (slice sl takes CTAs sl, sl + 8, ...: four loads in
// flight per thread), then a fixed-order sum of the 8 slices: one thread walking all n_cta partials of its output was a
// chain of ~150 dependent L2 loads = 29 us per launch whatever the size (profiles/r2p_launches_model_train.csv).
constexpr int RED_SLICES = 8, RED_OUT = 32;
__global__ void __launch_bounds__(RED_SLICES* RED_OUT)
    umma_tn_reduce_kernel(int n_cta, int n_q, int q_per_seg, int N, int64_t Kc, const float* __restrict__ partial,
                          float* __restrict__ C, int64_t s_kc, int64_t s_seg, int64_t s_n) {
  __shared__ float red[RED_SLICES][RED_OUT];
  const int64_t total = (int64_t)n_q * BLK * N;
  const int o = threadIdx.x % RED_OUT, sl = threadIdx.x / RED_OUT;
  for (int64_t e0 = (int64_t)blockIdx.x * RED_OUT; e0 < total; e0 += (int64_t)gridDim.x * RED_OUT) {
    const int64_t e = e0 + o;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (e < total) {
      int c = sl;
      for (; c + 3 * RED_SLICES < n_cta; c += 4 * RED_SLICES) {
        s0 += partial[(size_t)c * total + e];
        s1 += partial[(size_t)(c + RED_SLICES) * total + e];
        s2 += partial[(size_t)(c + 2 * RED_SLICES) * total + e];
        s3 += partial[(size_t)(c + 3 * RED_SLICES) * total + e];
      }
      for (; c < n_cta; c += RED_SLICES) s0 += partial[(size_t)c * total + e];
    }
    red[sl][o] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (sl == 0 && e < total) {
      float s = red[0][o];
#pragma unroll
      for (int j = 1; j < RED_SLICES; ++j) s += red[j][o];
      const int n = (int)(e % N);
      const int row = (int)(e / N);
      const int q = row / BLK, i = row % BLK;
      const int seg = q / q_per_seg;
      const int64_t kc = (int64_t)(q % q_per_seg) * BLK + i;
      if (kc < Kc) C[kc * s_kc + seg * s_seg + n * s_n] = s;
    }
    __syncthreads();
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn_tn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 3-D fp32 tensor [slabs, rows, cols] (contiguous), box = {32 cols, 16 rows, 1 slab}, SWIZZLE_128B_ATOM_32B;
// out-of-range rows / columns are zero-filled, which is what makes ragged R and Kc < 32 exact.
int make_rows_map(CUtensorMap* map, const float* base, int64_t slabs, int64_t rows, int64_t cols) {
  EncodeTiledFn fn = encode_fn_tn();
  DS_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)slabs};
  cuuint64_t strides[2] = {(cuuint64_t)cols * 4, (cuuint64_t)cols * 4 * (cuuint64_t)rows};
  cuuint32_t box[3] = {(cuuint32_t)BLK, (cuuint32_t)BKR, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DS_CHECK(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled (3-D) failed with code %d (slabs=%lld rows=%lld cols=%lld)",
           (int)rc, (long long)slabs, (long long)rows, (long long)cols);
  return 0;
}

struct TnGeometry {
  int q_per_seg, n_q, m_tiles, nb_blocks, stages;
  uint32_t stage_bytes, tmem_cols;
  size_t smem_bytes;
  int n_cta;
  int64_t rblocks_per_cta;
};

int tn_geometry(int64_t R, int64_t N, int64_t Kc, int nseg, int three, TnGeometry& g) {
  g.q_per_seg = (int)((Kc + BLK - 1) / BLK);
  g.n_q = nseg * g.q_per_seg;
  g.m_tiles = (g.n_q + 3) / 4;
  g.nb_blocks = (int)((N + BLK - 1) / BLK);
  uint32_t cols = 32;
  while ((int64_t)cols < (int64_t)g.m_tiles * N) cols <<= 1;
  if (cols > 512) return -1;
  g.tmem_cols = cols;
  // the last M tile may read up to 3 blocks past the A region: keep them inside the stage
  const uint32_t blocks = (uint32_t)(std::max(g.n_q, g.m_tiles * 4) + g.nb_blocks);
  g.stage_bytes = (three ? 2u : 1u) * (uint32_t)(g.n_q + g.nb_blocks) * BLK_BYTES;
  const uint32_t slack = (blocks - (uint32_t)(g.n_q + g.nb_blocks)) * BLK_BYTES;
  int dev = 0, max_smem = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return -1;
  const int ctl = (int)sizeof(TnCtl) + 1024 + (int)slack + 8192;
  int stages = (max_smem - ctl) / (int)g.stage_bytes;
  stages = std::min(stages, MAX_STAGES);
  if (stages < 2) return -1;
  g.stages = stages;
  g.smem_bytes = (size_t)stages * g.stage_bytes + ctl;
  const int64_t n_rblocks = (R + BKR - 1) / BKR;
  g.n_cta = (int)std::max<int64_t>(1, std::min<int64_t>(num_sms(), n_rblocks));
  g.rblocks_per_cta = (n_rblocks + g.n_cta - 1) / g.n_cta;
  return 0;
}

}  // namespace

int umma_tn_supported(int64_t R, int64_t N, int64_t Kc, int nseg) {
  if (Kc < 4 || Kc % 4 != 0 || N % 16 != 0 || N < 16 || N > 256) return -1;
  TnGeometry g;
  return tn_geometry(R, N, Kc, nseg, 1, g);
}

int64_t umma_tn_workspace_elems(int64_t R, int64_t N, int64_t Kc, int nseg) {
  TnGeometry g;
  if (tn_geometry(R, N, Kc, nseg, 1, g) != 0) return 0;
  return (int64_t)g.n_cta * g.n_q * BLK * N;
}

// A_seg: [R, Kc] contiguous (seg 0 = A0, others stacked in Arest with stride R rows); D: [R, N] contiguous
//   output element (kc, seg, n) goes to C[kc*s_kc + seg*s_seg + n*s_n]
int launch_umma_gemm_tn(int64_t R, int64_t N, int64_t Kc, int nseg, const float* A0, const float* Arest, const float* D,
                        float* C, int64_t s_kc, int64_t s_seg, int64_t s_n, float* partial, int mode, cudaStream_t st) {
  const int three = mode == DS_MODE_TF32X3 ? 1 : 0;
  TnGeometry g;
  DS_CHECK(tn_geometry(R, N, Kc, nseg, three, g) == 0, "umma tn: unsupported shape Kc=%lld N=%lld nseg=%d",
           (long long)Kc, (long long)N, nseg);
  TnParams p;
  p.R = R;
  p.N = (int)N;
  p.nb_blocks = g.nb_blocks;
  p.n_q = g.n_q;
  p.q_per_seg = g.q_per_seg;
  p.m_tiles = g.m_tiles;
  p.three_pass = three;
  p.stages = g.stages;
  p.stage_bytes = g.stage_bytes;
  p.tmem_cols = g.tmem_cols;
  p.rblocks_per_cta = g.rblocks_per_cta;
  p.partial = partial;
  p.A0 = A0; p.Arest = Arest != nullptr ? Arest : A0; p.D = D; p.Kc = Kc;
  static const int use_tma = [] { const char* e = getenv("DEEPSPHERE_TN_TMA"); return e && atoi(e) == 1; }();
  p.use_cp_async = (!use_tma && g.stages >= 5 && g.n_q + g.nb_blocks <= 16) ? 1 : 0;
  CUtensorMap m0, m1, md;
  DS_TRY(make_rows_map(&m0, A0, 1, R, Kc));
  DS_TRY(nseg > 1 ? make_rows_map(&m1, Arest, nseg - 1, R, Kc) : make_rows_map(&m1, A0, 1, R, Kc));
  DS_TRY(make_rows_map(&md, D, 1, R, N));
  static PerDeviceOnce attr_once;
  DS_TRY(attr_once.run([&]() -> int {
    int dev = 0, max_smem = 0;
    DS_CUDA(cudaGetDevice(&dev));
    DS_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    DS_CUDA(cudaFuncSetAttribute(umma_gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    return 0;
  }));
  umma_gemm_tn_kernel<<<g.n_cta, three ? 320 : 192, g.smem_bytes, st>>>(m0, m1, md, p);
  DS_LAUNCHED();
  const int64_t total = (int64_t)g.n_q * BLK * N;
  umma_tn_reduce_kernel<<<(unsigned)std::min<int64_t>((total + RED_OUT - 1) / RED_OUT, 4096), RED_SLICES * RED_OUT, 0, st>>>(
      g.n_cta, g.n_q, g.q_per_seg, (int)N, Kc, partial, C, s_kc, s_seg, s_n);
  DS_LAUNCHED();
  return 0;
}

}  // namespace ds
