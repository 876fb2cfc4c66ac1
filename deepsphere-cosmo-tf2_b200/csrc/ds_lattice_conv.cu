// Fully fused graph convolution on the HEALPix lattice (sm_100a): recursion + tensor-core contraction.
//
//   y[b, m, :] = act( sum_k T_k(L~) x [b, m, :] * W_k + bias )            (gnn_layers.py:131-159)
//
// Same tiling and stencil as ds_lattice.cu (16x16-pixel tile + H-ring halo in shared memory as float4
// channel planes, register-resident weights, sliding window).  What is new: the planes ARE the canonical
// no-swizzle K-major UMMA operand layout (8 consecutive lattice positions x 16 bytes = one core matrix,
// SBO = 128 B, LBO = plane stride), so after every hop ONE thread issues tcgen05.mma.kind::tf32 straight
// from the recursion's shared memory: A = T_k rows [H, H+16) of the lattice (384 positions = 3 M-tiles of
// 128, halo columns are computed and dropped), B = the 16-channel slice of W_k, accumulators in TMEM across
// all hops and channel chunks of one (tile, batch element).  No basis tensor ever goes to HBM and there is no
// separate GEMM kernel: HBM sees x (+halo, mostly L2) once and y once.
//
// 3xTF32 mode: the thread that produces a T_k value also stores its TF32 residual (v - trunc(v)) into a second
// operand tile; the issuer accumulates A_lo*B_hi + A_hi*B_lo + A_hi*B_hi (hardware truncates A itself).
//
// The same kernel serves the backward data pass: input dz, L~^T = L~, B = W_k^T, and it
// optionally writes the basis U_k = T_k(L~) dz (own pixels) that the weight-gradient kernel consumes.
#include <algorithm>
#include <cstdlib>

#include "ds_lattice.cuh"
#include "ds_ptx.cuh"

namespace ds {

struct LatConvArgs {
  int n_tiles;
  const int32_t* pix;
  const float* w;
  int64_t B, M;
  int F;        // input channels, F % 16 == 0
  int N;        // output channels, N % 16 == 0, 16 <= N <= 128
  int b_split;
  float alpha[LAT_MAX_STEPS], beta[LAT_MAX_STEPS];
  const float* in0;           // [B, M, F]
  float* out[LAT_MAX_STEPS];  // optional: basis of step s (own pixels), [B, M, F]
  const float* b_img;         // [F/16][H+1][parts][N*16] K-major no-swizzle images of the weight slices
  const float* bias;          // [N] or NULL
  int act;
  float* y;                   // [B, M, N]
};

namespace {

constexpr int CT = 16;   // tile side
constexpr int CFC = 16;  // channels per chunk
constexpr int CVPP = 4;  // float4 planes per chunk
constexpr int CS = 8;    // strip height

__host__ __device__ constexpr int conv_plane(int LW) {  // float4 per plane, == 2 (mod 8): one pad row above, the
  int v = (((LW + CS - 1) / CS) * CS + 3) * LW;       // strips' rows, two rows of look-ahead for the software pipeline
  while (v % 8 != 2) ++v;
  return v;
}
__host__ __device__ constexpr int conv_tasks(int LW) { return LW * ((LW + CS - 1) / CS); }
__host__ __device__ constexpr int conv_threads(int LW) { return ((conv_tasks(LW) * CVPP + 31) / 32) * 32; }
__host__ __device__ constexpr int conv_nld(int LW) { return (LW * LW * CVPP + conv_threads(LW) - 1) / conv_threads(LW); }
__host__ __device__ constexpr int conv_mrows(int LW) { return ((CT * LW + 127) / 128) * 128; }  // A rows (padded)

struct ConvCtl {
  uint64_t b_full[2];    // weight images of item parity landed
  uint64_t mma_done[2];  // all MMAs that read buffer parity p have completed
  uint32_t tmem_base;
};

__device__ __forceinline__ float4 c4_fma(float w, const float4& x, const float4& acc) {
  return make_float4(fmaf(w, x.x, acc.x), fmaf(w, x.y, acc.y), fmaf(w, x.z, acc.z), fmaf(w, x.w, acc.w));
}
__device__ __forceinline__ float4 c4_scale(float w, const float4& x) {
  return make_float4(w * x.x, w * x.y, w * x.z, w * x.w);
}
__device__ __forceinline__ float tf32_residual(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
__device__ __forceinline__ float4 c4_residual(const float4& v) {
  return make_float4(tf32_residual(v.x), tf32_residual(v.y), tf32_residual(v.z), tf32_residual(v.w));
}

template <int H, bool THREE>
__global__ void __launch_bounds__(conv_threads(CT + 2 * H), 1) lattice_conv_kernel(const LatConvArgs a) {
  constexpr int T = CT, LW = T + 2 * H, P = LW * LW, S = CS;
  constexpr int PL = conv_plane(LW);
  constexpr int TASKS = conv_tasks(LW), NT = conv_threads(LW);
  constexpr int N_LD = P * CVPP, NLD = conv_nld(LW);
  constexpr int MROWS = conv_mrows(LW), MT = MROWS / 128;  // operand rows: lattice rows [H, H+T), all columns
  constexpr int PLO = MROWS + 2;                           // residual-plane stride (float4), == 2 (mod 8)
  constexpr int PARTS = THREE ? 2 : 1;
  extern __shared__ __align__(128) uint8_t conv_smem[];
  float4* bufT[2];
  bufT[0] = reinterpret_cast<float4*>(conv_smem);
  bufT[1] = bufT[0] + CVPP * PL;
  float4* lo_buf[2];
  lo_buf[0] = bufT[1] + CVPP * PL + 8;  // + slack for the window wrap of the last plane
  lo_buf[1] = lo_buf[0] + (THREE ? CVPP * PLO : 0);
  const int N = a.N;
  const uint32_t img_bytes = (uint32_t)N * CFC * 4;  // one (chunk, hop, part) weight image
  const uint32_t bbuf_bytes = (uint32_t)(H + 1) * PARTS * img_bytes;
  uint8_t* bbuf[2];
  bbuf[0] = reinterpret_cast<uint8_t*>(lo_buf[1] + (THREE ? CVPP * PLO : 0));
  bbuf[1] = bbuf[0] + bbuf_bytes;
  int32_t* s_pix = reinterpret_cast<int32_t*>(bbuf[1] + bbuf_bytes);
  ConvCtl* ctl = reinterpret_cast<ConvCtl*>(s_pix + ((P + 3) & ~3));

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int task = tid % TASKS;
  const int fq = tid / TASKS;
  const int ci = task % LW;
  const int j0 = (task / LW) * S;
  const bool computes = tid < TASKS * CVPP;
  const int n_chunks = a.F / CFC;
  const int FV = a.F / 4;
  const int64_t b_per = (a.B + a.b_split - 1) / a.b_split;
  const int n_units = a.n_tiles * a.b_split;
  const int strip_off = fq * PL + (j0 + 1) * LW + ci;
  // residual-plane index of lattice position (j, i): rows [H, H+T) only
  const int lo_strip = fq * PLO + (j0 - H) * LW + ci;

  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < MT * N) tmem_cols <<= 1;
  if (tid == 0) {
    ptx::mbar_init(&ctl->b_full[0], 1);
    ptx::mbar_init(&ctl->b_full[1], 1);
    ptx::mbar_init(&ctl->mma_done[0], 1);
    ptx::mbar_init(&ctl->mma_done[1], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) ptx::tmem_alloc(&ctl->tmem_base, tmem_cols);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;
  const uint32_t idesc = ptx::make_idesc_tf32(128, N, 0, 0);

  // per-thread bookkeeping of barrier phases (control flow is uniform across the CTA)
  uint32_t n_commit[2] = {0, 0};  // commits issued so far on mma_done[p]
  uint32_t n_items_done = 0;      // items processed by this CTA (selects the weight-image buffer)
  auto wait_mma = [&](int p) {    // all MMAs reading buffer parity p are complete
    if (n_commit[p] > 0) ptx::mbar_wait(&ctl->mma_done[p], (n_commit[p] - 1) & 1);
  };

  struct LoadEvent { int64_t b; int c; bool valid; };
  float4 pre[NLD];
  auto issue = [&](const LoadEvent& ev) {
    if (!ev.valid) return;
    const float4* base = reinterpret_cast<const float4*>(a.in0 + (ev.b * a.M * a.F + ev.c * CFC));
#pragma unroll
    for (int r = 0; r < NLD; ++r) {
      const int u = tid + r * NT;
      pre[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (u < N_LD) {
        const int row = s_pix[u / CVPP];
        if (row >= 0) pre[r] = __ldg(base + (int64_t)row * FV + (u % CVPP));
      }
    }
  };

  for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const int tile = unit / a.b_split;
    const int64_t b_begin = (int64_t)(unit % a.b_split) * b_per;
    const int64_t b_end = min(a.B, b_begin + b_per);
    if (b_begin >= b_end) continue;
    // everything issued so far must be complete before the buffers are re-initialised
    wait_mma(0);
    wait_mma(1);
    __syncthreads();
    for (int p = tid; p < P; p += NT) s_pix[p] = a.pix[(size_t)tile * P + p];
    for (int e = tid; e < 2 * CVPP * PL + 8; e += NT) bufT[0][e] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (THREE)
      for (int e = tid; e < 2 * CVPP * PLO; e += NT) lo_buf[0][e] = make_float4(0.f, 0.f, 0.f, 0.f);
    float w[S][9];
#pragma unroll
    for (int jj = 0; jj < S; ++jj) {
      const int j = j0 + jj;
      const float* wp = a.w + ((size_t)tile * P + (size_t)min(j, LW - 1) * LW + ci) * 9;
#pragma unroll
      for (int d = 0; d < 9; ++d) w[jj][d] = (computes && j < LW) ? __ldg(wp + d) : 0.f;
    }
    __syncthreads();

    const int64_t n_items = (b_end - b_begin) * n_chunks;
    auto load_weights = [&](int64_t it) {  // weight images of item `it` -> bbuf[(n_items_done + it - cur) & 1]
      const int c = (int)(it % n_chunks);
      const uint32_t use = n_items_done + (uint32_t)it;  // global item counter of this CTA
      const int pb = use & 1;
      ptx::mbar_arrive_expect_tx(&ctl->b_full[pb], bbuf_bytes);
      ptx::bulk_load_1d(bbuf[pb], reinterpret_cast<const uint8_t*>(a.b_img) + (size_t)c * bbuf_bytes, bbuf_bytes,
                        &ctl->b_full[pb]);
    };
    // issue the UMMAs of hop k of the current item: A = T_k (buffer k & 1), B = image (chunk, k)
    auto issue_mma = [&](int k, uint32_t use, bool first_of_output) {
      const int pb = use & 1;
      if (k == 0) ptx::mbar_wait(&ctl->b_full[pb], (use >> 1) & 1);
      ptx::tc_fence_after_sync();
      const uint32_t a_base = ptx::smem_u32(bufT[k & 1] + LW + H * LW);  // plane 0, lattice row H, column 0
      const uint32_t lo_base = ptx::smem_u32(lo_buf[k & 1]);
      const uint32_t b_base = ptx::smem_u32(bbuf[pb]) + (uint32_t)k * PARTS * img_bytes;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int ks = 0; ks < CFC / 8; ++ks) {  // K = 8 per UMMA = 2 planes
          const uint64_t a_hi = ptx::make_smem_desc(a_base + (uint32_t)(ks * 2 * PL + mt * 128) * 16, PL * 16, 128,
                                                    ptx::LAYOUT_SWIZZLE_NONE);
          const uint64_t b_hi = ptx::make_smem_desc(b_base + (uint32_t)ks * 2 * N * 16, (uint32_t)N * 16, 128,
                                                    ptx::LAYOUT_SWIZZLE_NONE);
          const uint32_t d_tmem = tmem_base + (uint32_t)(mt * N);
          const uint32_t acc0 = (first_of_output && k == 0 && ks == 0) ? 0u : 1u;
          if (THREE) {
            const uint64_t a_lo = ptx::make_smem_desc(lo_base + (uint32_t)(ks * 2 * PLO + mt * 128) * 16, PLO * 16,
                                                      128, ptx::LAYOUT_SWIZZLE_NONE);
            const uint64_t b_lo = ptx::make_smem_desc(b_base + img_bytes + (uint32_t)ks * 2 * N * 16,
                                                      (uint32_t)N * 16, 128, ptx::LAYOUT_SWIZZLE_NONE);
            ptx::umma_tf32(d_tmem, a_lo, b_hi, idesc, acc0);
            ptx::umma_tf32(d_tmem, a_hi, b_lo, idesc, 1u);
            ptx::umma_tf32(d_tmem, a_hi, b_hi, idesc, 1u);
          } else {
            ptx::umma_tf32(d_tmem, a_hi, b_hi, idesc, acc0);
          }
        }
      }
      ptx::umma_commit(&ctl->mma_done[k & 1]);
    };

    if (tid == 0) load_weights(0);
    {
      LoadEvent first{b_begin, 0, true};
      issue(first);
    }
    for (int64_t it = 0; it < n_items; ++it) {
      const int64_t b = b_begin + it / n_chunks;
      const int c = (int)(it % n_chunks);
      const uint32_t use = n_items_done + (uint32_t)it;
      // ---- T_0: scatter the prefetched input (buffer 0), residuals for the operand rows ----
      wait_mma(0);
#pragma unroll
      for (int r = 0; r < NLD; ++r) {
        const int u = tid + r * NT;
        if (u < N_LD) {
          const int p = u / CVPP, q = u % CVPP;
          bufT[0][q * PL + LW + p] = pre[r];
          if (THREE) {
            const int m = p - H * LW;
            if (m >= 0 && m < T * LW) lo_buf[0][q * PLO + m] = c4_residual(pre[r]);
          }
        }
      }
      {
        LoadEvent nxt{b_begin + (it + 1) / n_chunks, (int)((it + 1) % n_chunks), it + 1 < n_items};
        issue(nxt);
      }
      ptx::fence_proxy_async_smem();
      __syncthreads();
      if (tid == 0) {
        issue_mma(0, use, c == 0);
      }
      n_commit[0]++;
      float4* cur = bufT[0];
      float4* oth = bufT[1];
#pragma unroll 1
      for (int s = 1; s <= H; ++s) {
        const float al = a.alpha[s - 1];
        const float be = a.beta[s - 1];
        const int pb = s & 1;
        wait_mma(pb);  // the UMMAs that read T_{s-2} (same buffer) are done: it may be overwritten
        if (s == 1 && tid == 0 && it + 1 < n_items) load_weights(it + 1);  // both barriers drained -> other B buffer free
        const int lo = s, hi = LW - 1 - s;
        if (computes && ci >= lo && ci <= hi && j0 <= hi && j0 + S - 1 >= lo) {
          const bool use_old = be != 0.f;
          const float4* cp = cur + strip_off;
          float4* op = oth + strip_off;
          float4* lp = lo_buf[pb] + lo_strip;
          float4 a0 = cp[-LW - 1], a1 = cp[-LW], a2 = cp[-LW + 1];
          float4 b0 = cp[-1], b1 = cp[0], b2 = cp[1];
          float4 c0 = cp[LW - 1], c1 = cp[LW], c2 = cp[LW + 1];
          float4 oldv = use_old ? op[0] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int jj = 0; jj < S; ++jj) {
            const float4 d0 = cp[(jj + 2) * LW - 1], d1 = cp[(jj + 2) * LW], d2 = cp[(jj + 2) * LW + 1];
            float4 oldn = make_float4(0.f, 0.f, 0.f, 0.f);
            if (use_old && jj + 1 < S) oldn = op[(jj + 1) * LW];
            float4 acc = c4_scale(w[jj][8], b1);
            acc = c4_fma(w[jj][0], b0, acc);
            acc = c4_fma(w[jj][1], c0, acc);
            acc = c4_fma(w[jj][2], c1, acc);
            acc = c4_fma(w[jj][3], c2, acc);
            acc = c4_fma(w[jj][4], b2, acc);
            acc = c4_fma(w[jj][5], a2, acc);
            acc = c4_fma(w[jj][6], a1, acc);
            acc = c4_fma(w[jj][7], a0, acc);
            float4 r = c4_scale(al, acc);
            if (use_old) r = c4_fma(be, oldv, r);
            const int j = j0 + jj;
            if (j >= lo && j <= hi) {
              op[jj * LW] = r;
              if (THREE && j >= H && j < H + T) lp[jj * LW] = c4_residual(r);
            }
            a0 = b0; a1 = b1; a2 = b2;
            b0 = c0; b1 = c1; b2 = c2;
            c0 = d0; c1 = d1; c2 = d2;
            oldv = oldn;
          }
        }
        ptx::fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) issue_mma(s, use, false);
        n_commit[pb]++;
        float* outp = a.out[s - 1];
        if (outp != nullptr) {
          float4* ob = reinterpret_cast<float4*>(outp + (b * a.M * a.F + c * CFC));
          for (int u = tid; u < T * T * CVPP; u += NT) {
            const int po = u / CVPP, q = u % CVPP;
            const int j = H + po / T, i = H + po % T;
            const int row = s_pix[j * LW + i];
            if (row >= 0) __stcs(ob + (int64_t)row * FV + q, oth[q * PL + (j + 1) * LW + i]);
          }
        }
        float4* t = cur; cur = oth; oth = t;
      }
      // ---- epilogue once per (tile, batch element): TMEM -> bias/activation -> y (own pixels) ----
      if (c == n_chunks - 1) {
        wait_mma(0);
        wait_mma(1);
        ptx::tc_fence_after_sync();
        if (warp < 4) {
          for (int mt = 0; mt < MT; ++mt) {
            const int m = mt * 128 + warp * 32 + lane;  // operand row = lattice position offset from row H
            const int j = H + m / LW, i = m % LW;
            int row = -1;
            if (m < T * LW && i >= H && i < H + T) row = s_pix[j * LW + i];
            float* yrow = a.y + (b * a.M + (row >= 0 ? row : 0)) * (int64_t)N;
            for (int c0 = 0; c0 < N; c0 += 16) {
              uint32_t r[16];
              ptx::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mt * N + c0), r);
              ptx::tmem_ld_wait();
              if (row >= 0) {
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                  float o[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    float val = __uint_as_float(r[v * 4 + e]);
                    if (a.bias != nullptr) val += __ldg(a.bias + c0 + v * 4 + e);
                    o[e] = act_apply(val, a.act);
                  }
                  __stcs(reinterpret_cast<float4*>(yrow + c0 + v * 4), make_float4(o[0], o[1], o[2], o[3]));
                }
              }
            }
          }
        }
        ptx::tc_fence_before_sync();
      }
      __syncthreads();  // basis stores / epilogue are done reading the buffers the next item overwrites
    }
    n_items_done += (uint32_t)n_items;
  }
  wait_mma(0);
  wait_mma(1);
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, tmem_cols);
  }
}

// weight images: for chunk c, hop k, part (hi | lo): element (n, kk) at (kk / 4) * (N * 4) + n * 4 + kk % 4
// value = Wsrc[(c*16 + kk) * s_f + k * s_k + n * s_n]
__global__ void conv_prep_b_kernel(const float* __restrict__ W, int64_t s_f, int64_t s_k, int64_t s_n, int n_chunks,
                                   int K, int N, int parts, float* __restrict__ img) {
  const int64_t per = (int64_t)N * CFC;
  const int64_t total = (int64_t)n_chunks * K * per;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int kk = (int)(e % CFC);
    const int n = (int)((e / CFC) % N);
    const int k = (int)((e / per) % K);
    const int c = (int)(e / (per * K));
    const float v = W[(int64_t)(c * CFC + kk) * s_f + (int64_t)k * s_k + (int64_t)n * s_n];
    uint32_t u = __float_as_uint(v);
    uint32_t rr = (u + 0x00000FFFu + ((u >> 13) & 1u)) & 0xFFFFE000u;
    float hi = __uint_as_float(rr);
    if (!isfinite(hi)) hi = __uint_as_float(u & 0xFFFFE000u);
    float* base = img + ((int64_t)(c * K + k) * parts) * per;
    const int64_t off = (int64_t)(kk / 4) * (N * 4) + n * 4 + (kk % 4);
    base[off] = hi;
    if (parts == 2) base[per + off] = v - hi;
  }
}

template <int H, bool THREE>
size_t conv_smem_bytes(int N) {
  constexpr int LW = CT + 2 * H;
  constexpr int PL = conv_plane(LW), MROWS = conv_mrows(LW), PLO = MROWS + 2;
  size_t b = (size_t)2 * CVPP * PL * 16 + 8 * 16;
  if (THREE) b += (size_t)2 * CVPP * PLO * 16;
  b += (size_t)2 * (H + 1) * (THREE ? 2 : 1) * N * CFC * 4;
  b += (size_t)((LW * LW + 3) & ~3) * 4 + sizeof(ConvCtl) + 128;
  return b;
}

template <int H, bool THREE>
int launch_conv_instance(const LatConvArgs& a, cudaStream_t st) {
  constexpr int LW = CT + 2 * H;
  const size_t smem = conv_smem_bytes<H, THREE>(a.N);
  int dev = 0, max_smem = 0;
  DS_CUDA(cudaGetDevice(&dev));
  DS_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  DS_CHECK((int)smem <= max_smem, "lattice conv: shared memory %zu > %d", smem, max_smem);
  static bool attr_done = false;
  if (!attr_done) {
    DS_CUDA(cudaFuncSetAttribute(lattice_conv_kernel<H, THREE>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    attr_done = true;
  }
  const int n_units = a.n_tiles * a.b_split;
  const int grid = std::min(n_units, num_sms());
  lattice_conv_kernel<H, THREE><<<grid, conv_threads(LW), smem, st>>>(a);
  DS_LAUNCHED();
  return 0;
}

template <int H>
bool conv_fits(int N, bool three) {
  int dev = 0, max_smem = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  if (cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return false;
  const size_t s = three ? conv_smem_bytes<H, true>(N) : conv_smem_bytes<H, false>(N);
  constexpr int LW = CT + 2 * H;
  return (int)s <= max_smem && conv_mrows(LW) / 128 * N <= 512;
}

}  // namespace

// Is the fully fused path available for this call? (H = K-1 in 1..4, F % 16 == 0, N % 16 == 0, N <= 128,
// tensor-core mode, shared memory and TMEM budget)
bool lattice_conv_usable(const LatticeDev& L, int F, int N, int mode) {
  if (mode == DS_MODE_FP32 || L.n_tiles <= 0 || L.T != CT) return false;
  if (F % 16 != 0 || N % 16 != 0 || N < 16 || N > 128) return false;
  // opt-in: this first fused kernel is shared-memory-bandwidth bound and slower than recursion + GEMM kernels;
  // the register-resident ds_lattice_conv2.cu supersedes it
  static const bool enabled = [] { const char* e = getenv("DEEPSPHERE_FUSED_CONV"); return e && atoi(e) == 1; }();
  if (!enabled) return false;
  const bool three = mode == DS_MODE_TF32X3;
  switch (L.H) {
    case 1: return conv_fits<1>(N, three);
    case 2: return conv_fits<2>(N, three);
    case 3: return conv_fits<3>(N, three);
    case 4: return conv_fits<4>(N, three);
    default: return false;
  }
}

// y = act( sum_k T_k(L~)(in0) * B_k + bias ) on the regular tiles; out[s-1] (optional) receives T_s(in0).
// Weights are addressed generically: B_k(f, n) = W[f*s_f + k*s_k + n*s_n].
int launch_lattice_conv(const LatticeDev& L, int64_t B, int64_t M, int F, int N, int recursion, const float* in0,
                        float* const* out, const float* W, int64_t s_f, int64_t s_k, int64_t s_n, const float* bias,
                        int act, float* y, int mode, cudaStream_t st) {
  const bool three = mode == DS_MODE_TF32X3;
  const int K = L.H + 1, n_chunks = F / CFC, parts = three ? 2 : 1;
  LatConvArgs a;
  a.n_tiles = L.n_tiles; a.pix = L.pix; a.w = L.w;
  a.B = B; a.M = M; a.F = F; a.N = N;
  int split = 1;
  while ((int64_t)L.n_tiles * split < (int64_t)8 * num_sms() && split < B) split *= 2;
  a.b_split = (int)std::min<int64_t>(split, B);
  for (int s = 0; s < LAT_MAX_STEPS; ++s) {
    const bool cheb2 = recursion == DS_RECURSION_CHEBYSHEV && s >= 1;  // step index s+1 >= 2
    a.alpha[s] = cheb2 ? 2.f : 1.f;
    a.beta[s] = cheb2 ? -1.f : 0.f;
    a.out[s] = (out != nullptr && s < L.H) ? out[s] : nullptr;
  }
  a.in0 = in0; a.bias = bias; a.act = act; a.y = y;
  float* img = nullptr;
  const size_t img_elems = (size_t)n_chunks * K * parts * N * CFC;
  DS_CUDA(cudaMallocAsync((void**)&img, img_elems * 4, st));
  conv_prep_b_kernel<<<(unsigned)std::min<size_t>((img_elems + 255) / 256, 1024), 256, 0, st>>>(W, s_f, s_k, s_n, n_chunks,
                                                                                               K, N, parts, img);
  g_launches.fetch_add(1);
  a.b_img = img;
  int rc = 0;
#define DS_CONV_CASE(HH) \
  case HH: rc = three ? launch_conv_instance<HH, true>(a, st) : launch_conv_instance<HH, false>(a, st); break;
  switch (L.H) {
    DS_CONV_CASE(1) DS_CONV_CASE(2) DS_CONV_CASE(3) DS_CONV_CASE(4)
    default: rc = fail("launch_lattice_conv: no instantiation for H=%d", L.H);
  }
#undef DS_CONV_CASE
  cudaFreeAsync(img, st);
  return rc;
}

}  // namespace ds
