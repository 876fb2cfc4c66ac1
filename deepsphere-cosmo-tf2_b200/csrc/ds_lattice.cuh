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


int lattice_configure(const LatticeDev& L, int64_t B, int64_t M, int F, LatticeArgs& a, int* threads, int* smem);
int launch_lattice(const LatticeDev& L, LatticeArgs& a, int threads, int smem, cudaStream_t st);

}  // namespace ds
