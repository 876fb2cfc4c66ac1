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

#include "ds_lattice.cuh"

namespace ds {

namespace {

constexpr int LAT_MAX_LD = 8;  // float4 prefetch registers per thread

// One "load event" = the lattice positions within ring limit `lo` (lo <= i, j <= LW-1-lo) of one
// [B, M, F] tensor for one (batch element, channel chunk), fetched into registers and later scattered
// into (or folded onto) a feature-major buffer.
struct LoadEvent {
  const float* src;
  int64_t b;
  int c, lo;
  bool valid;
};

__global__ void __launch_bounds__(320, 1) lattice_recursion_kernel(const LatticeArgs a) {
  extern __shared__ __align__(16) float lat_smem[];
  const int LW = a.LW, LWP = a.LWP, PS = a.PS, H = a.H, T = a.T, FC = a.FC;
  const int P = LW * LW;
  float* bufA = lat_smem;
  float* bufB = bufA + (size_t)FC * PS;
  int32_t* s_pix = reinterpret_cast<int32_t*>(bufB + (size_t)FC * PS);  // [P]

  const int tid = threadIdx.x, NT = blockDim.x;
  const int task = tid % a.tasks;
  const int fg = tid / a.tasks;
  const int ci = task % LW;            // lattice column of this thread's strip
  const int j0 = (task / LW) * LAT_S;  // first lattice row of the strip
  const bool computes = tid < a.tasks * a.nfg;
  const int n_chunks = a.F / FC;
  const int vpp = FC / 4;       // float4 per lattice position per chunk
  const int n_ld = P * vpp;     // float4 per load event (<= LAT_MAX_LD * NT, checked on the host)
  const int64_t b_per = (a.B + a.b_split - 1) / a.b_split;
  const int n_units = a.n_tiles * a.b_split;

  float4 pre[LAT_MAX_LD];
  auto issue = [&](const LoadEvent& ev) {
    if (!ev.valid) return;
#pragma unroll
    for (int r = 0; r < LAT_MAX_LD; ++r) {
      const int u = tid + r * NT;
      pre[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (u < n_ld) {
        const int p = u / vpp, q = u - p * vpp;
        const int j = p / LW, i = p - j * LW;
        const int row = s_pix[p];
        if (row >= 0 && i >= ev.lo && j >= ev.lo && i <= LW - 1 - ev.lo && j <= LW - 1 - ev.lo)
          pre[r] = __ldg(reinterpret_cast<const float4*>(ev.src + ((ev.b * a.M + row) * a.F + ev.c * FC)) + q);
      }
    }
  };
  // position (j, i) of plane f lives at f * PS + (j + 1) * LWP + (i + 1): one pad row / column around
  auto scatter = [&](float* dst) {  // dst = prefetched values (every lattice position)
#pragma unroll
    for (int r = 0; r < LAT_MAX_LD; ++r) {
      const int u = tid + r * NT;
      if (u < n_ld) {
        const int p = u / vpp, q = u - p * vpp;
        const int j = p / LW, i = p - j * LW;
        float* d = dst + (size_t)(4 * q) * PS + (j + 1) * LWP + (i + 1);
        d[0] = pre[r].x; d[PS] = pre[r].y; d[2 * PS] = pre[r].z; d[3 * PS] = pre[r].w;
      }
    }
  };
  auto fold = [&](float* dst, float be, float ga) {  // dst = be * dst + ga * prefetched
#pragma unroll
    for (int r = 0; r < LAT_MAX_LD; ++r) {
      const int u = tid + r * NT;
      if (u < n_ld) {
        const int p = u / vpp, q = u - p * vpp;
        const int j = p / LW, i = p - j * LW;
        float* d = dst + (size_t)(4 * q) * PS + (j + 1) * LWP + (i + 1);
        if (be == 0.f) {
          d[0] = ga * pre[r].x; d[PS] = ga * pre[r].y; d[2 * PS] = ga * pre[r].z; d[3 * PS] = ga * pre[r].w;
        } else {
          d[0] = fmaf(be, d[0], ga * pre[r].x); d[PS] = fmaf(be, d[PS], ga * pre[r].y);
          d[2 * PS] = fmaf(be, d[2 * PS], ga * pre[r].z); d[3 * PS] = fmaf(be, d[3 * PS], ga * pre[r].w);
        }
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
    for (int e = tid; e < 2 * FC * PS; e += NT) bufA[e] = 0.f;  // pads (and holes) stay zero for the whole tile
    // stencil weights of this thread's strip, register-resident for every item of the tile
    float w[LAT_S][9];
#pragma unroll
    for (int jj = 0; jj < LAT_S; ++jj) {
      const int j = j0 + jj;
      const float* wp = a.w + ((size_t)tile * P + (size_t)min(j, LW - 1) * LW + ci) * 9;
#pragma unroll
      for (int d = 0; d < 9; ++d) w[jj][d] = (computes && j < LW) ? __ldg(wp + d) : 0.f;
    }
    __syncthreads();

    const int64_t n_items = (b_end - b_begin) * n_chunks;
    // the event that follows (item it, step s_done): next add-input of the same item, else next item's input
    auto next_event = [&](int64_t it, int s_done) {
      LoadEvent ev;
      for (int s = s_done + 1; s <= a.nsteps; ++s)
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
      float* cur = bufA;
      float* oth = bufB;
      scatter(cur);            // consumes the prefetched input of this item ...
      issue(next_event(it, 0));  // ... and immediately puts the next event in flight
      __syncthreads();
      for (int s = 1; s <= a.nsteps; ++s) {
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
        if (computes && ci >= lo && ci <= hi) {
          for (int e = 0; e < a.fpt; ++e) {
            const int f = fg * a.fpt + e;
            const float* cp = cur + (size_t)f * PS + (j0 + 1) * LWP + (ci + 1);
            float* op = oth + (size_t)f * PS + (j0 + 1) * LWP + (ci + 1);
            float a0 = cp[-LWP - 1], a1 = cp[-LWP], a2 = cp[-LWP + 1];
            float b0 = cp[-1], b1 = cp[0], b2 = cp[1];
#pragma unroll
            for (int jj = 0; jj < LAT_S; ++jj) {
              const int j = j0 + jj;
              const float c0 = cp[(jj + 1) * LWP - 1], c1 = cp[(jj + 1) * LWP], c2 = cp[(jj + 1) * LWP + 1];
              if (j >= lo && j <= hi) {
                float acc = w[jj][8] * b1;
                acc = fmaf(w[jj][0], b0, acc);  // SW (-1, 0)
                acc = fmaf(w[jj][1], c0, acc);  // W  (-1,+1)
                acc = fmaf(w[jj][2], c1, acc);  // NW ( 0,+1)
                acc = fmaf(w[jj][3], c2, acc);  // N  (+1,+1)
                acc = fmaf(w[jj][4], b2, acc);  // NE (+1, 0)
                acc = fmaf(w[jj][5], a2, acc);  // E  (+1,-1)
                acc = fmaf(w[jj][6], a1, acc);  // SE ( 0,-1)
                acc = fmaf(w[jj][7], a0, acc);  // S  (-1,-1)
                const float oldv = be != 0.f ? be * op[jj * LWP] : 0.f;
                op[jj * LWP] = fmaf(al, acc, oldv);
              }
              a0 = b0; a1 = b1; a2 = b2;
              b0 = c0; b1 = c1; b2 = c2;
            }
          }
        }
        __syncthreads();
        // the step's result sits in `oth`: store the tile's own pixels if this step has an output
        float* outp = a.out[s - 1];
        if (outp != nullptr) {
          const int n_st = T * T * vpp;
          for (int u = tid; u < n_st; u += NT) {
            const int po = u / vpp, q = u - po * vpp;
            const int j = H + po / T, i = H + po % T;
            const int row = s_pix[j * LW + i];
            if (row >= 0) {
              const float* d = oth + (size_t)(4 * q) * PS + (j + 1) * LWP + (i + 1);
              __stcs(reinterpret_cast<float4*>(outp + ((b * a.M + row) * a.F + c * FC)) + q,
                     make_float4(d[0], d[PS], d[2 * PS], d[3 * PS]));
            }
          }
        }
        float* t = cur; cur = oth; oth = t;
      }
      __syncthreads();  // stores above read the buffers the next item overwrites
    }
  }
}

}  // namespace

int lattice_smem_bytes(int LW, int FC, int* LWP_out, int* PS_out) {
  int LWP = LW + 2;
  while (LWP % 4 != 3) ++LWP;  // 8 * LWP == 24 (mod 32): consecutive strips land on disjoint banks
  int PS = (LW + 2) * LWP;
  while (PS % 16 != 4) ++PS;   // feature-plane stride: 2 * PS == 8 (mod 32)
  if (LWP_out) *LWP_out = LWP;
  if (PS_out) *PS_out = PS;
  return (int)(2 * (size_t)FC * PS * 4 + (size_t)LW * LW * 4 + 64);
}

// choose the chunk width / thread layout; returns -1 if the lattice does not fit
int lattice_configure(const LatticeDev& L, int64_t B, int64_t M, int F, LatticeArgs& a, int* threads, int* smem) {
  if (L.n_tiles <= 0 || F % 4 != 0) return -1;
  int dev = 0, max_smem = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return -1;
  const int P = L.LW * L.LW;
  const int strips = (L.LW + LAT_S - 1) / LAT_S;
  const int tasks = L.LW * strips;
  if (tasks > 320) return -1;
  int FC = 16;
  while (FC > 4 && (F % FC != 0 || lattice_smem_bytes(L.LW, FC, nullptr, nullptr) > max_smem)) FC >>= 1;
  if (F % FC != 0 || lattice_smem_bytes(L.LW, FC, nullptr, nullptr) > max_smem) return -1;
  int nfg = std::min(FC, 320 / tasks);
  while (nfg > 1 && FC % nfg != 0) --nfg;
  const int nt = ((tasks * nfg + 31) / 32) * 32;
  if ((int64_t)P * (FC / 4) > (int64_t)8 * nt) return -1;  // prefetch registers (LAT_MAX_LD)
  a.n_tiles = L.n_tiles; a.LW = L.LW; a.H = L.H; a.T = L.T;
  a.pix = L.pix; a.w = L.w;
  a.B = B; a.M = M; a.F = F; a.FC = FC;
  *smem = lattice_smem_bytes(L.LW, FC, &a.LWP, &a.PS);
  a.tasks = tasks; a.nfg = nfg; a.fpt = FC / nfg;
  // enough work units to balance the SMs: split the batch when there are few tiles
  int split = 1;
  while ((int64_t)L.n_tiles * split < (int64_t)8 * num_sms() && split < B) split *= 2;
  a.b_split = (int)std::min<int64_t>(split, B);
  *threads = nt;
  return 0;
}

int launch_lattice(const LatticeDev& L, LatticeArgs& a, int threads, int smem, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    int dev = 0, max_smem = 0;
    DS_CUDA(cudaGetDevice(&dev));
    DS_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    DS_CUDA(cudaFuncSetAttribute(lattice_recursion_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    attr_done = true;
  }
  const int n_units = a.n_tiles * a.b_split;
  const int grid = std::min(n_units, num_sms());
  lattice_recursion_kernel<<<grid, threads, smem, st>>>(a);
  DS_LAUNCHED();
  return 0;
}

}  // namespace ds
