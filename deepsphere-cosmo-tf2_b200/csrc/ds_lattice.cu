// Fused multi-hop recursion on the HEALPix lattice (sm_100a).
//
// One CTA owns a tile (16 x 16 pixels of a base face) together with its H-ring halo, a (16+2H)^2 regular
// lattice kept in shared memory feature-major ([feature][row][column] planes).  All H = K-1 hops of
//      new = alpha_s * L~ cur + beta_s * old + gamma_s * add_s          (s = 1..H)
// run on chip: L~ acts as a 3x3 stencil with per-pixel weights that each thread keeps in REGISTERS for
// the strip of 8 pixels it owns (they are reused for every feature, batch element and hop), the data
// window slides down the strip so each hop costs ~4 shared-memory accesses per 9 FMAs, and the valid
// region shrinks by one ring per hop.  HBM sees each input once (plus the halo) instead of 3 tensors per
// hop.  This covers both directions of the path:
//   forward  (gnn_layers.py:135-143)  cur = x, no add, every hop's own-tile result stored (T_1..T_{K-1})
//   backward (Clenshaw adjoint)  cur = G_{K-1}, add_s = G_{K-1-s}, only the last hop stored (dx)
// Tiles whose neighbourhood is not a regular lattice (the 24 tiles at the valence-3 vertices) are not
// in the tile list; the caller runs the generic kernels on a compact sub-problem for them.
#include <algorithm>
#include <cstdlib>

#include "ds_lattice.cuh"

namespace ds {

namespace {

// ---- compile-time geometry of one (H, FC) instantiation ---------------------------------------------
// Shared-memory layout: float4 planes [FC/4][(LW + 2 + LAT_S) rows][LW columns]; a float4 holds 4 consecutive
// channels of one lattice position, so one 16-byte global load/store maps to one shared-memory access and the
// stencil runs on 4 channels per instruction.  Lattice position (j, i) sits at row j + 1 (one spare row above,
// LAT_S + 1 below for the last partial strip); there is no column padding: the i - 1 / i + 1 window of an edge
// column wraps into the neighbouring row, which only ever feeds the outermost ring that is never computed.
// With LW a multiple of 8 this is also the canonical no-swizzle K-major UMMA operand layout.
constexpr int LAT_T = 16;  // tile side
__host__ __device__ constexpr int lat_plane(int LW) {  // float4 per plane; == 2 (mod 8) so that the 4 channel
  int v = (LW + 2 + 8) * LW;  // groups of a position land on disjoint banks in the scatter / store phases
  while (v % 8 != 2) ++v;
  return v;
}
__host__ __device__ constexpr int lat_tasks(int LW, int S) { return LW * ((LW + S - 1) / S); }
__host__ __device__ constexpr int lat_threads(int LW, int FC, int S) {
  return ((lat_tasks(LW, S) * (FC / 4) + 31) / 32) * 32;
}
__host__ __device__ constexpr int lat_nld(int LW, int FC, int S) {  // float4 prefetch registers per thread
  return (LW * LW * (FC / 4) + lat_threads(LW, FC, S) - 1) / lat_threads(LW, FC, S);
}

struct LoadEvent {
  const float* src;
  int64_t b;
  int c, lo;
  bool valid;
};

__device__ __forceinline__ float4 f4_fma(float w, const float4& x, const float4& acc) {
  return make_float4(fmaf(w, x.x, acc.x), fmaf(w, x.y, acc.y), fmaf(w, x.z, acc.z), fmaf(w, x.w, acc.w));
}
__device__ __forceinline__ float4 f4_scale(float w, const float4& x) {
  return make_float4(w * x.x, w * x.y, w * x.z, w * x.w);
}

template <int H, int FC, int S>
__global__ void __launch_bounds__(lat_threads(LAT_T + 2 * H, FC, S), 1) lattice_recursion_kernel(const LatticeArgs a) {
  constexpr int T = LAT_T, LW = T + 2 * H, P = LW * LW;
  constexpr int PL = lat_plane(LW);
  constexpr int TASKS = lat_tasks(LW, S), VPP = FC / 4;
  constexpr int NT = lat_threads(LW, FC, S);
  constexpr int N_LD = P * VPP, NLD = lat_nld(LW, FC, S);
  extern __shared__ __align__(16) float4 lat_smem4[];
  float4* bufA = lat_smem4;
  float4* bufB = bufA + VPP * PL;
  int32_t* s_pix = reinterpret_cast<int32_t*>(bufB + VPP * PL);  // [P]

  const int tid = threadIdx.x;
  const int task = tid % TASKS;
  const int fq = tid / TASKS;          // which float4 channel group this thread computes
  const int ci = task % LW;            // lattice column of this thread's strip
  const int j0 = (task / LW) * S;      // first lattice row of the strip
  const bool computes = tid < TASKS * VPP;
  const int n_chunks = a.F / FC;
  const int FV = a.F / 4;
  const int64_t b_per = (a.B + a.b_split - 1) / a.b_split;
  const int n_units = a.n_tiles * a.b_split;
  const int strip_off = fq * PL + (j0 + 1) * LW + ci;

  float4 pre[NLD];
  auto issue = [&](const LoadEvent& ev) {
    if (!ev.valid) return;
    const float4* base = reinterpret_cast<const float4*>(ev.src + (ev.b * a.M * a.F + ev.c * FC));
#pragma unroll
    for (int r = 0; r < NLD; ++r) {
      const int u = tid + r * NT;
      pre[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (u < N_LD) {
        const int p = u / VPP, q = u % VPP;
        const int j = p / LW, i = p % LW;
        const int row = s_pix[p];
        if (row >= 0 && i >= ev.lo && j >= ev.lo && i <= LW - 1 - ev.lo && j <= LW - 1 - ev.lo)
          pre[r] = __ldg(base + (int64_t)row * FV + q);
      }
    }
  };
  auto scatter = [&](float4* dst) {
#pragma unroll
    for (int r = 0; r < NLD; ++r) {
      const int u = tid + r * NT;
      if (u < N_LD) {
        const int p = u / VPP, q = u % VPP;
        dst[q * PL + LW + p] = pre[r];  // (j + 1) * LW + i == LW + p
      }
    }
  };
  auto fold = [&](float4* dst, float be, float ga) {  // dst = be * dst + ga * prefetched
#pragma unroll
    for (int r = 0; r < NLD; ++r) {
      const int u = tid + r * NT;
      if (u < N_LD) {
        const int p = u / VPP, q = u % VPP;
        float4* d = dst + q * PL + LW + p;
        float4 v = f4_scale(ga, pre[r]);
        if (be != 0.f) v = f4_fma(be, *d, v);
        *d = v;
      }
    }
  };

  for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const int tile = unit / a.b_split;
    const int64_t b_begin = (int64_t)(unit % a.b_split) * b_per;
    const int64_t b_end = min(a.B, b_begin + b_per);
    if (b_begin >= b_end) continue;
    __syncthreads();  // previous unit is done with s_pix and the buffers
    for (int p = tid; p < P; p += NT) s_pix[p] = a.pix[(size_t)tile * P + p];
    for (int e = tid; e < 2 * VPP * PL; e += NT) bufA[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    // stencil weights of this thread's strip, register-resident for every item of the tile
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
    auto next_event = [&](int64_t it, int s_done) {
      LoadEvent ev;
      for (int s = s_done + 1; s <= H; ++s)
        if (a.add[s - 1] != nullptr) {
          ev.src = a.add[s - 1]; ev.b = b_begin + it / n_chunks; ev.c = (int)(it % n_chunks); ev.lo = s;
          ev.valid = true;
          return ev;
        }
      ev.valid = it + 1 < n_items;
      ev.src = a.in0; ev.b = b_begin + (it + 1) / n_chunks; ev.c = (int)((it + 1) % n_chunks); ev.lo = 0;
      return ev;
    };
    {
      LoadEvent first;
      first.src = a.in0; first.b = b_begin; first.c = 0; first.lo = 0; first.valid = true;
      issue(first);
    }
    for (int64_t it = 0; it < n_items; ++it) {
      const int64_t b = b_begin + it / n_chunks;
      const int c = (int)(it % n_chunks);
      float4* cur = bufA;
      float4* oth = bufB;
      scatter(cur);              // consumes the prefetched input of this item ...
      issue(next_event(it, 0));  // ... and immediately puts the next event in flight
      __syncthreads();
#pragma unroll 1
      for (int s = 1; s <= H && s <= a.nsteps; ++s) {
        const float al = a.alpha[s - 1];
        float be = a.beta[s - 1];
        if (a.add[s - 1] != nullptr) {
          // old' = beta * old + gamma * add, then the stencil pass adds alpha * L~ cur on top
          fold(oth, be, a.gamma[s - 1]);
          issue(next_event(it, s));
          be = 1.f;
          __syncthreads();
        }
        const int lo = s, hi = LW - 1 - s;  // region computed by this step
        if (computes && ci >= lo && ci <= hi && j0 <= hi && j0 + S - 1 >= lo) {
          const bool use_old = be != 0.f;
          const float4* cp = cur + strip_off;
          float4* op = oth + strip_off;
          // explicit software pipeline: the window row needed by iteration jj + 1 and the `old` value of
          // iteration jj + 1 are requested one iteration ahead, so ~40 FMAs cover each shared-memory round trip
          float4 a0 = cp[-LW - 1], a1 = cp[-LW], a2 = cp[-LW + 1];
          float4 b0 = cp[-1], b1 = cp[0], b2 = cp[1];
          float4 c0 = cp[LW - 1], c1 = cp[LW], c2 = cp[LW + 1];
          float4 oldv = use_old ? op[0] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int jj = 0; jj < S; ++jj) {
            const float4 d0 = cp[(jj + 2) * LW - 1], d1 = cp[(jj + 2) * LW], d2 = cp[(jj + 2) * LW + 1];
            float4 oldn = make_float4(0.f, 0.f, 0.f, 0.f);
            if (use_old && jj + 1 < S) oldn = op[(jj + 1) * LW];
            float4 acc = f4_scale(w[jj][8], b1);
            acc = f4_fma(w[jj][0], b0, acc);  // SW (-1, 0)
            acc = f4_fma(w[jj][1], c0, acc);  // W  (-1,+1)
            acc = f4_fma(w[jj][2], c1, acc);  // NW ( 0,+1)
            acc = f4_fma(w[jj][3], c2, acc);  // N  (+1,+1)
            acc = f4_fma(w[jj][4], b2, acc);  // NE (+1, 0)
            acc = f4_fma(w[jj][5], a2, acc);  // E  (+1,-1)
            acc = f4_fma(w[jj][6], a1, acc);  // SE ( 0,-1)
            acc = f4_fma(w[jj][7], a0, acc);  // S  (-1,-1)
            float4 r = f4_scale(al, acc);
            if (use_old) r = f4_fma(be, oldv, r);
            const int j = j0 + jj;
            if (j >= lo && j <= hi) op[jj * LW] = r;
            a0 = b0; a1 = b1; a2 = b2;
            b0 = c0; b1 = c1; b2 = c2;
            c0 = d0; c1 = d1; c2 = d2;
            oldv = oldn;
          }
        }
        __syncthreads();
        // the step's result sits in `oth`: store the tile's own pixels if this step has an output
        float* outp = a.out[s - 1];
        if (outp != nullptr) {
          float4* ob = reinterpret_cast<float4*>(outp + (b * a.M * a.F + c * FC));
          for (int u = tid; u < T * T * VPP; u += NT) {
            const int po = u / VPP, q = u % VPP;
            const int j = H + po / T, i = H + po % T;
            const int row = s_pix[j * LW + i];
            if (row >= 0) __stcs(ob + (int64_t)row * FV + q, oth[q * PL + (j + 1) * LW + i]);
          }
        }
        float4* t = cur; cur = oth; oth = t;
      }
      __syncthreads();  // stores above read the buffers the next item overwrites
    }
  }
}

template <int H, int FC, int S>
int launch_instance(const LatticeArgs& a, cudaStream_t st) {
  constexpr int LW = LAT_T + 2 * H;
  constexpr int smem = 2 * (FC / 4) * lat_plane(LW) * 16 + LW * LW * 4 + 64;
  static std::atomic<int> ctas_per_sm{0};  // identical on every device of one box
  static PerDeviceOnce attr_once;
  DS_TRY(attr_once.run([&]() -> int {
    DS_CUDA(cudaFuncSetAttribute(lattice_recursion_kernel<H, FC, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int n = 1;
    DS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, lattice_recursion_kernel<H, FC, S>, lat_threads(LW, FC, S), smem));
    ctas_per_sm = std::max(1, n);
    return 0;
  }));
  const int n_units = a.n_tiles * a.b_split;
  const int grid = std::min(n_units, num_sms() * ctas_per_sm);
  lattice_recursion_kernel<H, FC, S><<<grid, lat_threads(LW, FC, S), smem, st>>>(a);
  DS_LAUNCHED();
  return 0;
}

}  // namespace

// instantiated geometries: H = K - 1 in 1..6, channel chunk 16 (F % 16 == 0) or 4
int lattice_configure(const LatticeDev& L, int64_t B, int64_t M, int F, LatticeArgs& a, int* threads, int* smem) {
  if (L.n_tiles <= 0 || F % 4 != 0 || L.T != LAT_T || L.H < 1 || L.H > 6 || L.LW != LAT_T + 2 * L.H) return -1;
  a.n_tiles = L.n_tiles; a.LW = L.LW; a.H = L.H; a.T = L.T;
  a.pix = L.pix; a.w = L.w;
  a.B = B; a.M = M; a.F = F; a.FC = F % 16 == 0 ? 16 : 4;
  a.LWP = 0; a.PS = 0; a.tasks = 0; a.nfg = 0; a.fpt = 0;
  int split = 1;  // enough work units to balance the SMs: split the batch when there are few tiles
  while ((int64_t)L.n_tiles * split < (int64_t)8 * num_sms() && split < B) split *= 2;
  a.b_split = (int)std::min<int64_t>(split, B);
  *threads = 0;
  *smem = 0;
  return 0;
}

int launch_lattice(const LatticeDev& L, LatticeArgs& a, int threads, int smem, cudaStream_t st) {
  (void)L; (void)threads; (void)smem;
  static const bool strip4 = [] { const char* e = getenv("DEEPSPHERE_LATTICE_STRIP"); return e && atoi(e) == 4; }();
#define DS_LAT_CASE(HH)                                                        \
  case HH:                                                                     \
    return a.FC == 16 ? (strip4 ? launch_instance<HH, 16, 4>(a, st) : launch_instance<HH, 16, 8>(a, st)) \
                      : launch_instance<HH, 4, 8>(a, st);
  switch (a.H) {
    DS_LAT_CASE(1) DS_LAT_CASE(2) DS_LAT_CASE(3) DS_LAT_CASE(4) DS_LAT_CASE(5) DS_LAT_CASE(6)
    default: break;
  }
#undef DS_LAT_CASE
  return fail("launch_lattice: no instantiation for H=%d", a.H);
}

}  // namespace ds
