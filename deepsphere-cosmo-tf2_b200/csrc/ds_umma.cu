// tcgen05 tensor-core contraction for DS_MODE_TF32 / DS_MODE_TF32X3:
//   C[R, N] = act( A_cat[R, Ktot] * Bt[N, Ktot]^T + bias ),   A_cat = [A_0 | A_1 | ... | A_{nseg-1}]
// This is tf.matmul(x_stack, kernel) of gnn_layers.py:149 with the K basis tensors T_k[B*M, Fin]
// read in place as the K-segments of the A operand (no tf.stack / transposes, :144-147), and
// equally its data gradient dz * W_k^T.
//
// Structure (one persistent CTA per SM, warp-specialised, sm_100a only):
//   warp 0      TMA producer: per 32-column k-block one cp.async.bulk.tensor (A: 128 rows x 128 B,
//               SWIZZLE_128B, straight from the [R, Kc] tensor) + one bulk copy of the pre-swizzled
//               B image, completing on the stage's `full` mbarrier
//   warp 1      UMMA issuer: one elected thread issues tcgen05.mma.kind::tf32 (M128 x N x K8) on
//               shared-memory descriptors, accumulating in TMEM (two accumulators, so the epilogue of
//               tile i overlaps the MMAs of tile i+1); tcgen05.commit frees the stage
//   warps 2-5   epilogue: tcgen05.ld the accumulator (lane = row), bias + activation, 16-byte stores
//   warps 6-9   (3xTF32 only) operand splitter: A = A_hi + A_lo with A_hi = fp32 truncated to TF32
//               precision; the issuer then accumulates A_lo*B_hi + A_hi*B_lo + A_hi*B_hi
// Roofline: HBM-bound (reads (nseg+.)*R*Kc*4 bytes, writes R*N*4); tensor pipe has >2x slack even at 3x.
#include <algorithm>
#include <mutex>

#include "ds_common.cuh"
#include "ds_ptx.cuh"

namespace ds {
namespace {

constexpr int BM = 128;   // rows per tile = UMMA M
constexpr int BK = 32;    // fp32 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 8; // K of one tcgen05.mma.kind::tf32
constexpr int A_TILE_BYTES = BM * BK * 4;  // 16 KB
constexpr int NUM_THREADS_TF32 = 6 * 32;
constexpr int NUM_THREADS_3X = 10 * 32;
constexpr int MAX_STAGES = 8;

struct UmmaParams {
  int64_t R;            // rows of A and C
  int N;                // columns of C (UMMA N)
  int nseg;             // K segments
  int kb_per_seg;       // ceil(Kc / 32)
  int64_t rest_rows;    // row offset between consecutive segments inside the `rest` tensor map (= R)
  const float* b_img;   // [n_kb][parts][N][32] pre-swizzled; parts = 1 (hi) or 2 (hi, lo)
  const float* bias;    // nullable
  int bias_mod;
  int act;
  float* C;
  int64_t ldc;
  int three_pass;
  int stages;
  uint32_t stage_bytes;
  uint32_t tmem_cols;
};

// B image: for k-block kb = seg*kb_per_seg + q and column n, element j (k = q*32 + j):
//   K-major SWIZZLE_128B row n of 128 bytes, 16-byte chunk index XOR (n % 8)
__global__ void umma_prep_b_kernel(const float* __restrict__ Bm, int64_t ldb, int64_t k_stride, int64_t seg_stride,
                                   int64_t n_stride, int64_t kcol_stride, int N, int Kc, int nseg, int kb_per_seg,
                                   int parts, float* __restrict__ img) {
  // B(k = (seg, kc), n) = Bm[(kc*k_stride + seg*seg_stride + n*n_stride) * ldb + kc*kcol_stride + n*(1-...)]
  // expressed through two generic strides so both the NN form (rows = k, cols = n) and the NT form
  // (rows = n, cols = k) are covered:
  //   element offset = kc * k_stride + seg * seg_stride + n * n_stride          (all in elements)
  (void)ldb; (void)kcol_stride;
  const int n_kb = nseg * kb_per_seg;
  const int64_t total = (int64_t)n_kb * N * BK;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(e % BK);
    const int n = (int)((e / BK) % N);
    const int kb = (int)(e / ((int64_t)BK * N));
    const int seg = kb / kb_per_seg, q = kb % kb_per_seg;
    const int kc = q * BK + j;
    float v = 0.f;
    if (kc < Kc) v = Bm[(int64_t)kc * k_stride + (int64_t)seg * seg_stride + (int64_t)n * n_stride];
    // round-to-nearest TF32 split: hi keeps 10 explicit mantissa bits, lo is the exact remainder
    uint32_t u = __float_as_uint(v);
    uint32_t r = (u + 0x00000FFFu + ((u >> 13) & 1u)) & 0xFFFFE000u;
    float hi = __uint_as_float(r);
    if (!isfinite(hi)) hi = __uint_as_float(u & 0xFFFFE000u);
    const float lo = v - hi;
    const int chunk = (j >> 2) ^ (n & 7);
    const int64_t off = (int64_t)n * BK + chunk * 4 + (j & 3);
    float* base = img + (int64_t)kb * parts * N * BK;
    base[off] = hi;
    if (parts == 2) base[(int64_t)N * BK + off] = lo;
  }
}

struct SmemLayout {
  // dynamic smem: [stages x stage_bytes] aligned to 1024, then barriers
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t split_done[MAX_STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

// TMEM -> registers (lane = row) -> bias/activation -> swizzled shared-memory tile -> one TMA store per
// 32-column block: full 128-byte lines leave the SM instead of 16-byte pieces at a 256-byte stride.
// The staging buffer holds 64 columns; wider outputs are flushed in 64-column groups.
template <int ACT>
__device__ __noinline__ void epilogue_loop(const UmmaParams& p, SmemLayout* ctl, uint8_t* c_stage,
                                           const CUtensorMap* map_c, uint32_t tmem_base, int64_t n_tiles) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3;               // TMEM lane quarter this warp may access
  const int et = threadIdx.x - 2 * 32;  // 0..127 among the epilogue threads
  const int r_in_tile = q * 32 + lane;
  const bool has_bias = p.bias != nullptr;
  uint32_t tile_it = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
    const uint32_t acc = tile_it & 1;
    const uint32_t acc_ph = (tile_it >> 1) & 1;
    ptx::mbar_wait(&ctl->tmem_full[acc], acc_ph);
    ptx::tc_fence_after_sync();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)p.N;
    for (int g0 = 0; g0 < p.N; g0 += 64) {
      // the previous TMA stores must have finished reading the staging buffer
      if (et == 0) ptx::tma_store_wait_read();
      ptx::named_bar_sync(1, 128);
      const int g1 = min(p.N, g0 + 64);
      for (int c0 = g0; c0 < g1; c0 += 16) {
        uint32_t r[16];
        ptx::tmem_ld_32x32b_x16(taddr + c0, r);
        ptx::tmem_ld_wait();
        if (c0 + 16 >= p.N) {
          // accumulator fully read: hand it back to the issuer before the math / stores
          ptx::tc_fence_before_sync();
          ptx::mbar_arrive(&ctl->tmem_empty[acc]);
        }
        uint8_t* blk = c_stage + (size_t)((c0 - g0) >> 5) * A_TILE_BYTES + (size_t)r_in_tile * 128;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float val = __uint_as_float(r[v * 4 + e]);
            if (has_bias) val += __ldg(p.bias + ((c0 + v * 4 + e) % p.bias_mod));
            o[e] = act_apply(val, ACT);
          }
          const int chunk = (((c0 & 31) >> 2) + v) ^ (r_in_tile & 7);  // 16-byte chunk, 128B swizzle
          *reinterpret_cast<float4*>(blk + chunk * 16) = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::named_bar_sync(1, 128);
      if (et == 0) {
        for (int cb = 0; cb < (g1 - g0 + 31) / 32; ++cb)
          ptx::tma_store_2d(map_c, c_stage + (size_t)cb * A_TILE_BYTES, g0 + cb * 32, (int32_t)(tile * BM));
        ptx::tma_store_commit();
      }
    }
  }
  if (et == 0) ptx::tma_store_wait_all();
}

__global__ void __launch_bounds__(NUM_THREADS_3X, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_arest,
                 const __grid_constant__ CUtensorMap map_c, const UmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // stage buffers first (1024-byte aligned for SWIZZLE_128B), control block after them
  uint8_t* stage_base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // epilogue staging: up to 2 blocks of [128 rows x 128 B], SWIZZLE_128B (what the C tensor map expects)
  uint8_t* c_stage = stage_base + (size_t)p.stages * p.stage_bytes;
  SmemLayout* ctl = reinterpret_cast<SmemLayout*>(c_stage + (size_t)min((p.N + 31) / 32, 2) * A_TILE_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_kb = p.nseg * p.kb_per_seg;
  const int64_t n_tiles = (p.R + BM - 1) / BM;
  const uint32_t b_part_bytes = (uint32_t)p.N * BK * 4;
  const uint32_t parts = p.three_pass ? 2u : 1u;
  // stage layout: [A_hi 16K][A_lo 16K (3x)][B_hi][B_lo (3x)]
  const uint32_t off_alo = A_TILE_BYTES;
  const uint32_t off_b = p.three_pass ? 2 * A_TILE_BYTES : A_TILE_BYTES;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&ctl->full[s], 1);
      ptx::mbar_init(&ctl->empty[s], 1);
      ptx::mbar_init(&ctl->split_done[s], 128);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&ctl->tmem_full[a], 1);
      ptx::mbar_init(&ctl->tmem_empty[a], 128);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tma_prefetch_desc(&map_a0);
    ptx::tma_prefetch_desc(&map_arest);
    ptx::tma_prefetch_desc(&map_c);
  }
  if (warp == 1) ptx::tmem_alloc(&ctl->tmem_base, p.tmem_cols);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (ptx::elect_one()) {
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int32_t row0 = (int32_t)(tile * BM);
        for (int kb = 0; kb < n_kb; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          ptx::mbar_wait(&ctl->empty[s], ph ^ 1);
          uint8_t* st = stage_base + (size_t)s * p.stage_bytes;
          ptx::mbar_arrive_expect_tx(&ctl->full[s], A_TILE_BYTES + parts * b_part_bytes);
          const int seg = kb / p.kb_per_seg, q = kb % p.kb_per_seg;
          if (seg == 0) ptx::tma_load_2d(st, &map_a0, q * BK, row0, &ctl->full[s]);
          else ptx::tma_load_2d(st, &map_arest, q * BK, (int32_t)((seg - 1) * p.rest_rows) + row0, &ctl->full[s]);
          ptx::bulk_load_1d(st + off_b, p.b_img + (size_t)kb * parts * p.N * BK, parts * b_part_bytes, &ctl->full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== UMMA issuer =====================
    const uint32_t idesc = ptx::make_idesc_tf32(BM, p.N, 0, 0);
    uint32_t it = 0, tile_it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
      const uint32_t acc = tile_it & 1;
      const uint32_t acc_ph = (tile_it >> 1) & 1;
      ptx::mbar_wait(&ctl->tmem_empty[acc], acc_ph ^ 1);
      ptx::tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.N;
      for (int kb = 0; kb < n_kb; ++kb, ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1;
        ptx::mbar_wait(&ctl->full[s], ph);
        if (p.three_pass) ptx::mbar_wait(&ctl->split_done[s], ph);
        ptx::tc_fence_after_sync();
        if (ptx::elect_one()) {
          const uint32_t st = ptx::smem_u32(stage_base + (size_t)s * p.stage_bytes);
          const uint64_t a_hi = ptx::make_smem_desc(st, 0, 1024, ptx::LAYOUT_SWIZZLE_128B);
          const uint64_t a_lo = ptx::make_smem_desc(st + off_alo, 0, 1024, ptx::LAYOUT_SWIZZLE_128B);
          const uint64_t b_hi = ptx::make_smem_desc(st + off_b, 0, 1024, ptx::LAYOUT_SWIZZLE_128B);
          const uint64_t b_lo = ptx::make_smem_desc(st + off_b + b_part_bytes, 0, 1024, ptx::LAYOUT_SWIZZLE_128B);
#pragma unroll
          for (int j = 0; j < BK / UMMA_K; ++j) {
            const uint64_t adv = (uint64_t)((j * UMMA_K * 4) >> 4);  // +32 bytes per K step inside the swizzle row
            const uint32_t first = (kb == 0 && j == 0) ? 0u : 1u;
            if (p.three_pass) {
              ptx::umma_tf32(d_tmem, a_lo + adv, b_hi + adv, idesc, first);
              ptx::umma_tf32(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
              ptx::umma_tf32(d_tmem, a_hi + adv, b_hi + adv, idesc, 1u);
            } else {
              ptx::umma_tf32(d_tmem, a_hi + adv, b_hi + adv, idesc, first);
            }
          }
          ptx::umma_commit(&ctl->empty[s]);  // frees the smem stage once these MMAs have read it
          if (kb == n_kb - 1) ptx::umma_commit(&ctl->tmem_full[acc]);
        }
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ===================== epilogue (4 warps = 128 TMEM lanes) =====================
    // instantiated per activation: one inlined 6-way switch per output value made the epilogue ~43 KB of
    // SASS and instruction-fetch bound (ncu r1e: stall_no_instruction 5.0 per issue)
    switch (p.act) {
      case DS_ACT_RELU: epilogue_loop<DS_ACT_RELU>(p, ctl, c_stage, &map_c, tmem_base, n_tiles); break;
      case DS_ACT_ELU: epilogue_loop<DS_ACT_ELU>(p, ctl, c_stage, &map_c, tmem_base, n_tiles); break;
      case DS_ACT_SIGMOID: epilogue_loop<DS_ACT_SIGMOID>(p, ctl, c_stage, &map_c, tmem_base, n_tiles); break;
      case DS_ACT_TANH: epilogue_loop<DS_ACT_TANH>(p, ctl, c_stage, &map_c, tmem_base, n_tiles); break;
      case DS_ACT_SOFTPLUS: epilogue_loop<DS_ACT_SOFTPLUS>(p, ctl, c_stage, &map_c, tmem_base, n_tiles); break;
      default: epilogue_loop<DS_ACT_LINEAR>(p, ctl, c_stage, &map_c, tmem_base, n_tiles); break;
    }
  } else if (p.three_pass) {
    // ===================== 3xTF32 operand splitter (4 warps) =====================
    const int t = threadIdx.x - 6 * 32;  // 0..127
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < n_kb; ++kb, ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1;
        ptx::mbar_wait(&ctl->full[s], ph);
        float4* hi = reinterpret_cast<float4*>(stage_base + (size_t)s * p.stage_bytes);
        float4* lo = reinterpret_cast<float4*>(stage_base + (size_t)s * p.stage_bytes + off_alo);
#pragma unroll
        for (int i = 0; i < A_TILE_BYTES / 16 / 128; ++i) {
          const int idx = i * 128 + t;
          const float4 v = hi[idx];
          float4 h, l;
          h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
          h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
          h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
          h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
          // the tensor core truncates fp32 operands to TF32 itself (probed: tools/umma_probe.py), so the
          // high part needs no write-back: only the residual tile is materialised
          lo[idx] = l;
        }
        ptx::fence_proxy_async_smem();  // generic-proxy writes -> visible to the UMMA operand fetch
        ptx::mbar_arrive(&ctl->split_done[s]);
      }
    }
  }

  // teardown: everyone done with TMEM before the allocating warp frees it
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ---- host side --------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
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

// 2-D fp32 tensor [rows, cols] (row stride = ld elements), box = 128 rows x 32 columns, SWIZZLE_128B
// (used for the A operand loads and for the C tile stores)
int make_a_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld) {
  EncodeTiledFn fn = encode_fn();
  DS_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DS_CHECK(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d (rows=%lld cols=%lld ld=%lld)", (int)rc,
           (long long)rows, (long long)cols, (long long)ld);
  return 0;
}

uint32_t pow2_cols(int n) {
  uint32_t c = 32;
  while ((int)c < n) c <<= 1;
  return c;
}

}  // namespace

// Kc % 4 == 0 (16-byte row pitch for TMA), N % 16 == 0, 16 <= N <= 256
int umma_supported(int64_t Kc, int nseg, int64_t N) {
  (void)nseg;
  if (Kc < 4 || Kc % 4 != 0) return -1;
  if (N % 16 != 0 || N < 16 || N > 256) return -1;
  return 0;
}

// C[R,N] = act( sum_seg A_seg[R,Kc] * B(seg,kc,n) + bias[col % bias_mod] ), A rows contiguous (lda = Kc).
//   B(seg,kc,n) = Bm[kc*b_k_stride + seg*b_seg_stride + n*b_n_stride]   (element strides)
int launch_umma_gemm(int64_t R, int64_t N, int64_t Kc, int nseg, const float* A0, const float* Arest,
                     int64_t a_seg_stride_rows, const float* Bm, int64_t b_k_stride, int64_t b_seg_stride,
                     int64_t b_n_stride, const float* bias, int64_t bias_mod, int act, float* C, int64_t ldc, int mode,
                     cudaStream_t st) {
  DS_CHECK(umma_supported(Kc, nseg, N) == 0, "umma gemm: unsupported shape Kc=%lld N=%lld", (long long)Kc, (long long)N);
  DS_CHECK(nseg == 1 || a_seg_stride_rows == R, "umma gemm: segments must be stacked with stride R rows");
  DS_CHECK((reinterpret_cast<uintptr_t>(A0) & 15) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0 && ldc % 4 == 0,
           "umma gemm: operands must be 16-byte aligned");
  const int three = mode == DS_MODE_TF32X3 ? 1 : 0;
  UmmaParams p;
  p.R = R;
  p.N = (int)N;
  p.nseg = nseg;
  p.kb_per_seg = (int)((Kc + BK - 1) / BK);
  p.rest_rows = R;
  p.bias = bias;
  p.bias_mod = (int)(bias_mod > 0 ? bias_mod : 1);
  p.act = act;
  p.C = C;
  p.ldc = ldc;
  p.three_pass = three;
  const int parts = three ? 2 : 1;
  p.stage_bytes = (uint32_t)((three ? 2 : 1) * A_TILE_BYTES + parts * N * BK * 4);
  p.tmem_cols = pow2_cols((int)(2 * N));
  int dev = 0, max_smem = 0;
  DS_CUDA(cudaGetDevice(&dev));
  DS_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const int ctl_bytes = (int)sizeof(SmemLayout) + 1024 /* alignment slack */ + (int)std::min<int64_t>((N + 31) / 32, 2) * A_TILE_BYTES;
  int stages = (max_smem - ctl_bytes) / (int)p.stage_bytes;
  stages = std::min(stages, std::min(MAX_STAGES, three ? 4 : 6));
  DS_CHECK(stages >= 2, "umma gemm: tile does not fit shared memory (stage %u bytes)", p.stage_bytes);
  p.stages = stages;
  const size_t smem_bytes = (size_t)stages * p.stage_bytes + ctl_bytes;

  // B image (pre-swizzled, hi/lo split) in a stream-ordered scratch allocation
  const int n_kb = nseg * p.kb_per_seg;
  const size_t img_bytes = (size_t)n_kb * parts * N * BK * 4;
  float* img = nullptr;
  DS_CUDA(cudaMallocAsync((void**)&img, img_bytes, st));
  {
    const int64_t total = (int64_t)n_kb * N * BK;
    umma_prep_b_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, 1024), 256, 0, st>>>(
        Bm, 0, b_k_stride, b_seg_stride, b_n_stride, 0, (int)N, (int)Kc, nseg, p.kb_per_seg, parts, img);
    g_launches.fetch_add(1);
  }
  p.b_img = img;

  CUtensorMap map0, map1, mapc;
  int rc = make_a_map(&mapc, C, R, N, ldc);
  if (rc == 0) rc = make_a_map(&map0, A0, R, Kc, Kc);
  if (rc == 0) rc = nseg > 1 ? make_a_map(&map1, Arest, (int64_t)(nseg - 1) * R, Kc, Kc) : make_a_map(&map1, A0, R, Kc, Kc);
  if (rc != 0) {
    cudaFreeAsync(img, st);
    return rc;
  }
  static PerDeviceOnce attr_once;
  rc = attr_once.run([&]() -> int {
    DS_CUDA(cudaFuncSetAttribute(umma_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    return 0;
  });
  if (rc != 0) {
    cudaFreeAsync(img, st);
    return rc;
  }
  const int64_t n_tiles = (R + BM - 1) / BM;
  const unsigned grid = (unsigned)std::min<int64_t>(n_tiles, num_sms());
  umma_gemm_kernel<<<grid, three ? NUM_THREADS_3X : NUM_THREADS_TF32, smem_bytes, st>>>(map0, map1, mapc, p);
  g_launches.fetch_add(1);
  cudaError_t le = cudaGetLastError();
  cudaFreeAsync(img, st);
  DS_CHECK(le == cudaSuccess, "umma_gemm_kernel launch failed: %s", cudaGetErrorString(le));
  return 0;
}

// the forward contraction form used by ds_graph_conv_forward: B(seg,kc,n) = Bm[(kc*bks + seg*bss) * N + n]
int launch_umma_gemm_nn(int64_t R, int64_t N, int64_t Kc, int nseg, const float* A0, const float* Arest,
                        int64_t a_seg_stride, const float* Bm, int64_t b_kc_stride, int64_t b_seg_stride,
                        const float* bias, int act, float* C, int mode, cudaStream_t st) {
  return launch_umma_gemm(R, N, Kc, nseg, A0, Arest, a_seg_stride / Kc, Bm, b_kc_stride * N, b_seg_stride * N, 1, bias,
                          N, act, C, N, mode, st);
}

}  // namespace ds
