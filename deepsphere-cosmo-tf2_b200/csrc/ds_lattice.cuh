// Declarations shared by the lattice kernel (ds_lattice.cu) and its C-ABI glue (ds_lattice_api.cu).
#pragma once

#include "ds_common.cuh"

namespace ds {

constexpr int LAT_MAX_STEPS = 16;
constexpr int LAT_S = 8;  // strip height (pixels per thread per feature)

struct LatticeDev {
  int n_tiles = 0;   // tiles in the tables (all that have an exact own pixel)
  int LW = 0, H = 0, T = 0;
  int32_t* pix = nullptr;  // [n_tiles][LW*LW]
  float* w = nullptr;      // [n_tiles][LW*LW][9]
  // centre weight (the diagonal of L~) of every lattice position that carries a pixel: ONE value when the plan comes from a
  // normalised Laplacian (L~_ii = 2*scale/lmax - 1) - the round-2 fused kernel keeps it as a scalar
  bool diag_const = false;
  float diag = 0.f;
};

struct LatticeArgs {
  int n_tiles, LW, H, T;
  const int32_t* pix;
  const float* w;
  int64_t B, M;
  int F, FC;             // channels, channels per chunk (F % FC == 0, FC % 4 == 0)
  int LWP, PS;           // padded plane row stride / plane stride (floats)
  int tasks, nfg, fpt;   // (column, strip) tasks, feature groups, features per thread
  int b_split;           // work units per tile along the batch
  int nsteps;
  const float* in0;                   // initial `cur`, [B, M, F]
  const float* add[LAT_MAX_STEPS];    // optional per-step additive input
  float* out[LAT_MAX_STEPS];          // optional per-step output (own pixels)
  float alpha[LAT_MAX_STEPS], beta[LAT_MAX_STEPS], gamma[LAT_MAX_STEPS];
};


// the irregular rows of the plan grouped into connected patches of their H-hop closure (ds_patch.cu)
struct PatchDev {
  int n_patches = 0, max_rows = 0, max_own = 0;
  int32_t* row_ptr = nullptr;    // [n_patches + 1] ranges of `rows`
  int32_t* rows = nullptr;       // global row of every patch row
  int32_t* ell_col = nullptr;    // [n_rows][9] patch-local column (-1: none)
  float* ell_val = nullptr;      // [n_rows][9]
  int32_t* own_ptr = nullptr;    // [n_patches + 1] ranges of `own_local`
  int32_t* own_local = nullptr;  // patch-local rows whose result is wanted
};
bool patch_usable(const PatchDev& P, int nsteps, int F, int N);
int launch_patch_conv(const PatchDev& P, int nsteps, int64_t B, int64_t M, int F, int N, int recursion, const float* in0,
                      float* const* out, const float* W, int64_t s_f, int64_t s_k, int64_t s_n, const float* bias,
                      int act, float* y, cudaStream_t st);

int lattice_configure(const LatticeDev& L, int64_t B, int64_t M, int F, LatticeArgs& a, int* threads, int* smem);
int launch_lattice(const LatticeDev& L, LatticeArgs& a, int threads, int smem, cudaStream_t st);

}  // namespace ds
