// Register-resident fused graph convolution on the HEALPix lattice (sm_100a), warp specialised.
//
//   y[b, m, :] = act( sum_k T_k(L~) x [b, m, :] * W_k + bias )            (gnn_layers.py:131-159)
//
// Work item = (16x16-pixel tile + 4-ring halo = 24x24 lattice, batch element b, 8-channel chunk c).
// What differs from ds_lattice.cu and from the first fused kernel (git history: ds_lattice_conv.cu), whose recursion is
// shared-memory-bandwidth bound at ~5 accesses per 36 FMAs: every compute thread OWNS a 3x3 pixel block x 4 channels for all hops of an item and keeps
// T_{k-1} and T_{k-2} of that block in REGISTERS next to its 81 stencil weights.  Shared memory only carries
// the neighbour exchange: per hop a thread stores its 9 new values and loads the 16 perimeter values
// (2.8 accesses per 36 FMAs), and the very same exchange buffer is the K-major no-swizzle UMMA A operand
// (8 channels = one tf32 K step), so the contraction costs no extra shared-memory writes.
//
// Roles (256 threads, 2 CTAs / SM so that one CTA's synchronisation phases overlap the other's arithmetic;
// registers rebalanced with setmaxnreg, see C2_REG_*):
//   warps 0-3  compute: 64 blocks x 2 channel quads; lane = (block-row half, quad q, column block cb).  The arithmetic
//              is packed FFMA2 with the weight as broadcast scalar operand.  Hops are synchronised with an mbarrier
//              (every thread arrives after its stores + proxy fence) instead of __syncthreads, and the taps that
//              only need the thread's own registers (49 of 81) run BEFORE the wait for the neighbours.  These warps
//              also drain the TMEM accumulator (bias + activation + store of the own pixels) of the previous (tile, b).
//   warp  4    one lane: streams the weight slice of the next chunk (cp.async.bulk) and issues
//              tcgen05.mma.kind::tf32 after every hop: A = T_k rows 4..19 of the lattice (384 positions = 3 M-tiles),
//              B = 8-channel slice of W_k, accumulating over hops and chunks in TMEM.
//   warps 5-7  gather the next item's input rows into a staging buffer (cp.async, zero fill for holes).
//
// Exchange-buffer layout: 2 float4 planes (channel quads) of 26 x 24 positions; lattice (row j, column c) sits
// at position (j + 1) * 24 + (c % 3) * 8 + c / 3, i.e. the three columns of a block are de-interleaved so that
// the 8 lanes of a quarter-warp always touch 8 consecutive float4 (conflict-free LDS.128 / STS.128), and 8
// consecutive positions are one UMMA core matrix (SBO = 128 B, LBO = plane stride).
//
// Chebyshev: the register-resident weights are 2 L~ (exact doubling), so T_k = (2L~) T_{k-1} - T_{k-2} costs exactly
// 9 FMAs per output (the subtraction rides on the first FMA); hop 1 halves its result (T_1 = L~ T_0).
// Monomial: plain T_k = L~ T_{k-1}.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "ds_lattice.cuh"
#include "ds_ptx.cuh"

namespace ds {

namespace {

constexpr int C2_LW = 24, C2_T = 16, C2_H = 4, C2_P = C2_LW * C2_LW;
constexpr int C2_FC = 8;               // channels per work item = one UMMA K step (tf32: K = 8)
constexpr int C2_NQ = C2_FC / 4;       // float4 planes per buffer
constexpr int C2_PL = 26 * C2_LW + 4;  // float4 per plane (+4: the two planes start on complementary bank groups)
constexpr int C2_BUF = C2_NQ * C2_PL;  // float4 per exchange buffer
// Rows of the pixel block a compute thread owns (3 columns wide).  3 (default, the measured kernel): 64 blocks x 2
// channel quads = 4 compute warps, 81 weights + 72 basis values per thread.  2 (experiment, -DC2_BR=2): 96 blocks =
// 6 compute warps, 54 weights + 48 basis values per thread -> 136 registers, i.e. 12 instead of 8 compute warps per
// SM to hide the per-hop dependency chain, for 14 instead of 16 perimeter loads per 54 instead of 81 FMAs.
#ifndef C2_BR
#define C2_BR 3
#endif
static_assert(C2_BR == 3 || C2_BR == 2, "block rows");
// Channels per compute thread.  4 (round 1): a thread owns a float4 of its block's pixels, 4 compute warps per CTA.
// 2 (round 2): a thread owns HALF a float4 (one FFMA2 per tap), two neighbouring lanes share a position: 8 compute
// warps per CTA = 16 per SM against the per-hop dependency chain (exchange -> barrier -> perimeter loads), half the
// basis values and perimeter registers per thread (no spills), the same shared-memory bytes per FMA.
#ifndef C2_CPT
#define C2_CPT 4
#endif
static_assert(C2_CPT == 4 || (C2_CPT == 2 && C2_BR == 3), "channels per thread");
// Columns of the pixel block (C2_BR rows).  3 (round 1).  6 (round 2, with C2_CPT = 2): 3 x 6 pixels x 2 channels per thread,
// the same 72 basis registers and 4 compute warps as the 3 x 3 x 4 shape, but 22 perimeter positions per 18 pixels
// instead of 16 per 9: the exchange loads - the largest share of the shared-memory wavefronts, which the ncu counters
// show to be the kernel's ceiling (60 % busy at 31 % FMA) - drop by a third, and 112 of the 162 FFMA2 of a hop only
// need the thread's own registers (they run before the wait for the neighbours).
#ifndef C2_BC
#define C2_BC 3
#endif
static_assert(C2_BC == 3 || (C2_BC == 6 && C2_CPT == 2 && C2_BR == 3), "block columns");
constexpr int C2_NB = C2_LW / C2_BC;  // blocks per lattice row; the columns of a block are de-interleaved (see slot_of_col)
constexpr int C2_VP = 4 / C2_CPT;  // vectors of a thread's width per 16-byte position
constexpr int C2_NCW = (C2_LW / C2_BR) * C2_NB * 2 * C2_VP / 32;  // compute warps (one thread per block, quad and half)
// position of lattice column c inside its row: the C2_NB blocks' columns cc sit next to each other, so that the lanes
// of a quarter- / half-warp touch consecutive 16-byte positions; 8 consecutive positions are one UMMA core matrix
__host__ __device__ constexpr int slot_of_col(int c) { return (c % C2_BC) * C2_NB + c / C2_BC; }
__host__ __device__ constexpr int col_of_slot(int p) { return (p % C2_NB) * C2_BC + p / C2_NB; }
#ifndef C2_NLOAD
#define C2_NLOAD 96
#endif
static_assert(C2_NLOAD % 32 == 0 && (2 * C2_LW * C2_LW) % C2_NLOAD == 0, "gather threads");
// 1 = the accumulator drain (TMEM -> bias / activation -> y) runs on the IO warps instead of the compute warps: the three
// gather warps (CTA warps 5, 6, 7 = TMEM lane quarters 1, 2, 3) between their gathers, and the UMMA-issuing warp (CTA warp
// 4 = quarter 0) right after it has committed a group's last UMMAs.  ncu source page of the round-2 default build: 17.7 %
// of the compute warps' samples sit in the drain, while the IO warps sleep 82 % of the time.  (A fifth IO warp for quarter
// 0 does not work: setmaxnreg is a WARPGROUP instruction - all four warps of an aligned group of four must execute the
// same one, which is also why the 6-compute-warp variants of round 1 fault.)
#ifndef C2_IO_DRAIN
#define C2_IO_DRAIN 0
#endif
static_assert(!C2_IO_DRAIN || (C2_NCW == 4 && C2_NLOAD == 96), "IO-side drain: warps 4..7 must cover the four TMEM lane quarters");
// 1 = the per-hop basis stores of a launch with `out` (the backward-data launch hands U_k = T_k(L~) dz to the weight-gradient
// kernel: 4A bytes; forward 14.3 ms, the same launch with the stores 20.5 ms) are done by the GATHER warps from the exchange
// buffer - every hop's T_k of all positions is in shared memory anyway - instead of by the compute threads from registers
// (9 row look-ups, 9 address computations and 9 scattered 16-byte stores per thread and hop in the hop's dependency chain).
#ifndef C2_IO_OUT
#define C2_IO_OUT 0
#endif

constexpr int C2_NCOMP = 32 * C2_NCW, C2_THREADS = C2_NCOMP + 32 + C2_NLOAD;
// setmaxnreg is a WARPGROUP instruction: the four warps of an aligned group of four must all execute the same one.  The
// compute warps (inc) therefore have to fill whole warpgroups - the 6-warp shape of C2_BR = 2 (round 1, host-emulated only)
// faults on the B200 (round 2, call 1) - and the IO warps (dec) are exactly one warpgroup.
#ifndef DS_EMULATE
static_assert(C2_NCW % 4 == 0 && C2_NLOAD == 96, "compute and IO warps must each fill whole warpgroups (setmaxnreg)");
#endif
constexpr int C2_NEPI = C2_NCW >= 8 ? 8 : 4;  // warps that drain the accumulators (TMEM lane quarter = warp % 4)
// registers per CTA (two CTAs per SM): the pool is what the launch allocates, threads x (registers per thread of the
// launch bound, a multiple of 8); setmaxnreg moves it between the roles.  Overridable for the variant builds.
constexpr int C2_REG_POOL = (32768 / C2_THREADS / 8) * 8 * C2_THREADS;
#ifndef C2_REGS_IO
#define C2_REGS_IO (C2_CPT == 2 ? 32 : 48)
#endif
#ifndef C2_REGS_COMPUTE
#define C2_REGS_COMPUTE (((C2_REG_POOL - (C2_THREADS - C2_NCOMP) * C2_REGS_IO) / C2_NCOMP) / 8 * 8)
#endif
constexpr int C2_REG_COMPUTE = C2_REGS_COMPUTE, C2_REG_IO = C2_REGS_IO;
#ifndef C2_SKIP_REG_ASSERT
static_assert(C2_NCOMP * C2_REG_COMPUTE + (C2_THREADS - C2_NCOMP) * C2_REG_IO <= C2_REG_POOL, "register budget of two CTAs per SM");
#endif
constexpr int C2_TMEM_COLS = 256;

struct Conv2Args {
  int n_tiles;
  const int32_t* pix;
  const float* w;
  int64_t B, M;
  int F, N;
  int b_split;
  int nsteps;  // hops, 1..4
  float wscale;
  float wdiag;              // C2_CDIAG: the (already scaled) centre weight of every pixel
  long long* dbg;           // optional timeline probe [item < 16][hop 0..4][8 slots] of clock64 (block 0)
  int sleep_mma, sleep_ld;  // ns of back-off between polls of the issuer / gather roles (0: plain spin)
  const float* in0;     // [B, M, F]
  float* out[C2_H];     // optional basis of hop s (own pixels), [B, M, F]
  const float* b_img;   // [F/8][nsteps+1][N*8] K-major no-swizzle images of the weight slices
  const float* bias;    // [N] or NULL
  int act;
  float* y;             // [B, M, N]
};

struct Conv2Ctl {
  uint64_t in_full[2], in_empty[2], w_full[2], item_done[2], hop_full[2], mma_done[2], acc_full, acc_empty;
#if C2_IO_OUT
  uint64_t out_done[2];  // the gather warps have copied the hop in X[p] to global memory (launches with `out`)
#endif
#if C2_SPLIT_BAR
  uint64_t hop_ready[2];
#endif
  uint32_t tmem_base;
};

// Register-bank-aware arithmetic.  A float4 that goes through LDS.128 / STS.128 sits in 4 consecutive, 4-aligned
// registers, so member i of EVERY such vector has the same register parity; a scalar FFMA reading `in.x` and
// `acc.x` takes two reads from one of the two register banks and issues at half rate (measured: 2.1 clk / FFMA,
// and ptxas re-orders the taps so the weight rarely sits in the operand reuse cache).  The packed FFMA2 form reads
// aligned register PAIRS (one register from each bank per operand), halves the issued instructions and cannot be
// split by the scheduler; the weight is broadcast into a register pair.
#ifndef C2_FFMA2
#define C2_FFMA2 1
#endif
// Experiment switch (round 2, A/B through DEEPSPHERE_LIB): 1 = the generic->async proxy fence that orders the hop's
// exchange stores before the UMMA reads is executed ONCE by the issuing lane after it has observed hop_full, instead of
// by each of the 128 compute threads before its arrive (~300 clk on every hop's dependency chain).  Whether the
// consumer-side fence is sufficient is exactly what the parity tests of the variant build have to show; default 0.
#ifndef C2_FENCE_BY_ISSUER
#define C2_FENCE_BY_ISSUER 0
#endif
// ROT = 0: `x` natural channel order, `acc` rotated by one (x.x <-> acc.w, x.y <-> acc.x, ...); ROT = 1: `x` rotated,
// `acc` natural; ROT = 2: both natural (used with FFMA2).
__host__ __device__ constexpr int rot_of_hop(int s) { return C2_FFMA2 ? 2 : ((s - 1) & 1); }
__host__ __device__ constexpr bool hop_is_rotated(int s) { return !C2_FFMA2 && (s & 1); }
template <int ROT>
__device__ __forceinline__ void f4_fma(float w, const float4& x, float4& acc) {
  if (ROT == 2) {
    const float2 ww = make_float2(w, w);
    const float2 lo = __ffma2_rn(ww, make_float2(x.x, x.y), make_float2(acc.x, acc.y));
    const float2 hi = __ffma2_rn(ww, make_float2(x.z, x.w), make_float2(acc.z, acc.w));
    acc = make_float4(lo.x, lo.y, hi.x, hi.y);
  } else if (ROT == 0) {
    acc.w = fmaf(w, x.x, acc.w); acc.x = fmaf(w, x.y, acc.x); acc.y = fmaf(w, x.z, acc.y); acc.z = fmaf(w, x.w, acc.z);
  } else {
    acc.y = fmaf(w, x.x, acc.y); acc.z = fmaf(w, x.y, acc.z); acc.w = fmaf(w, x.z, acc.w); acc.x = fmaf(w, x.w, acc.x);
  }
}
template <int ROT>
__device__ __forceinline__ void f4_fms(float w, const float4& x, float4& acc) {  // acc = w*x - acc
  if (ROT == 2) {
    const float2 ww = make_float2(w, w);
    const float2 lo = __ffma2_rn(ww, make_float2(x.x, x.y), make_float2(-acc.x, -acc.y));
    const float2 hi = __ffma2_rn(ww, make_float2(x.z, x.w), make_float2(-acc.z, -acc.w));
    acc = make_float4(lo.x, lo.y, hi.x, hi.y);
  } else if (ROT == 0) {
    acc.w = fmaf(w, x.x, -acc.w); acc.x = fmaf(w, x.y, -acc.x); acc.y = fmaf(w, x.z, -acc.y); acc.z = fmaf(w, x.w, -acc.z);
  } else {
    acc.y = fmaf(w, x.x, -acc.y); acc.z = fmaf(w, x.y, -acc.z); acc.w = fmaf(w, x.z, -acc.w); acc.x = fmaf(w, x.w, -acc.x);
  }
}
template <int ROT>
__device__ __forceinline__ void f4_mul(float w, const float4& x, float4& acc) {  // acc = w*x
  if (ROT == 2) {
    acc = make_float4(w * x.x, w * x.y, w * x.z, w * x.w);
  } else if (ROT == 0) {
    acc.w = w * x.x; acc.x = w * x.y; acc.y = w * x.z; acc.z = w * x.w;
  } else {
    acc.y = w * x.x; acc.z = w * x.y; acc.w = w * x.z; acc.x = w * x.w;
  }
}
__device__ __forceinline__ float4 f4_scale(float w, const float4& x) {
  return make_float4(w * x.x, w * x.y, w * x.z, w * x.w);
}
// two channels per thread (C2_CPT == 2): one packed FFMA2 per tap
template <int ROT>
__device__ __forceinline__ void f4_fma(float w, const float2& x, float2& acc) {
  acc = __ffma2_rn(make_float2(w, w), x, acc);
}
template <int ROT>
__device__ __forceinline__ void f4_fms(float w, const float2& x, float2& acc) {
  acc = __ffma2_rn(make_float2(w, w), x, make_float2(-acc.x, -acc.y));
}
template <int ROT>
__device__ __forceinline__ void f4_mul(float w, const float2& x, float2& acc) {
  acc = make_float2(w * x.x, w * x.y);
}
__device__ __forceinline__ float2 f4_scale(float w, const float2& x) { return make_float2(w * x.x, w * x.y); }
#if C2_CPT == 2
typedef float2 cvec;
static_assert(C2_FFMA2, "two channels per thread need the packed form");
#else
typedef float4 cvec;
#endif
__device__ __noinline__ float4 act4(float4 v, int act) {
  return make_float4(act_apply(v.x, act), act_apply(v.y, act), act_apply(v.z, act), act_apply(v.w, act));
}
#ifndef DS_EMULATE
__device__ __forceinline__ void cp_async16(void* dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(ptx::smem_u32(dst)), "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#define C2_SETMAXNREG_INC(n) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(n))
#define C2_SETMAXNREG_DEC(n) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(n))
#define C2_DYNAMIC_SMEM(name) extern __shared__ __align__(128) uint8_t name[]
// volatile: keeps the scaled weight in its register (ptxas otherwise re-multiplies at every use)
#define C2_SCALED_WEIGHT(dst, v, scale) asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(dst) : "f"(v), "f"(scale))
#else  // host emulation (tests/emul): same semantics with plain C++
inline void cp_async16(void* dst, const void* src, uint32_t src_bytes) {
  std::memset(dst, 0, 16);
  std::memcpy(dst, src, src_bytes);
}
inline void cp_async_wait_all() {}
#define C2_SETMAXNREG_INC(n) ((void)0)
#define C2_SETMAXNREG_DEC(n) ((void)0)
#define C2_DYNAMIC_SMEM(name) uint8_t* const name = emul::dynamic_smem()
#define C2_SCALED_WEIGHT(dst, v, scale) (dst) = (v) * (scale)
#endif

// stencil direction index of the plan's weight table for (drow, dcol) (lattice.py: SW, W, NW, N, NE, E, SE, S, centre)
__device__ __forceinline__ constexpr int dir_of(int dr, int dc) {
  return dr == 0 ? (dc < 0 ? 0 : (dc > 0 ? 4 : 8)) : (dr > 0 ? (dc < 0 ? 1 : (dc == 0 ? 2 : 3)) : (dc > 0 ? 5 : (dc == 0 ? 6 : 7)));
}

// Experiment switch: 1 = exploit L~_pq = L~_qp for the links BETWEEN two pixels of the thread's own block: the weight is
// kept once (by the pixel that comes first in the block), 20 of the 81 weight registers of a 3x3 block and 11 of the 54
// of a 2x3 block disappear.  L~ is symmetric up to fp32 rounding of the two stored copies (1e-6 relative, the tolerance
// the plan already accepts for using the same stencil in the backward pass), so results move at that level.
#ifndef C2_SYMW
#define C2_SYMW 0
#endif
// 0 = compile the clock64 timeline probe (DEEPSPHERE_CONV2_DEBUG) out of the kernel: its predicates, branches and
// constant loads are ~5 % of the instructions the compute warps issue (ncu source page, r1k capture)
#ifndef C2_PROBE
#define C2_PROBE 0  // round 2 default (measured r2c: 15.5 -> 14.3 ms forward); build with -DC2_PROBE=1 for the timeline
#endif
// Experiment switch: 1 = two barriers per hop.  hop_ready: every compute thread arrives right after its exchange stores
// (generic proxy; mbarrier arrive / wait are release / acquire) and it is what the NEXT hop's perimeter loads wait for;
// the generic->async proxy fence and the arrive on hop_full (what the UMMA issuer waits for) move behind the next hop's
// own-register taps, where the stores have long landed: the ~300 clk of FENCE.VIEW.ASYNC leave the hop-to-hop
// dependency chain without touching the ordering the UMMA reads rely on (each thread still fences its own stores
// before it arrives on hop_full).
#ifndef C2_SPLIT_BAR
#define C2_SPLIT_BAR 0
#endif
// Experiment switch: 1 = software-pipelined accumulator drain (see the epilogue)
#ifndef C2_EPI_PIPE
#define C2_EPI_PIPE 0
#endif
// 1 = the centre weight (the diagonal of L~) is ONE scalar for the whole plan (normalised Laplacians: a - 1, SURVEY F8;
// checked on the host when the lattice is attached, Conv2Args::wdiag): 9 (18) weight registers per thread less
#ifndef C2_CDIAG
#define C2_CDIAG (C2_BC == 6)
#endif
// 1 = the warp that owns the outermost block rows (lattice rows 0-2 and 21-23) sits out the hops whose valid region no
// longer reaches them (the last two hops of an item): no loads, FMAs or stores, only the barrier arrivals
#ifndef C2_SKIP_OUTER
#define C2_SKIP_OUTER 0  // measured (r2b): 15.3 ms with, 14.6 ms without - the divergence costs more than the idle warp saves
#endif
// 1 = the hops of an item run as a LOOP over one code body per buffer parity instead of four unrolled, individually
// specialised hops.  Why: the ncu captures show the GPC-level instruction cache at 84 - 90 % of its peak request rate
// (gcc__cache_requests_type_instruction; the SM's own instruction cache misses 10 % of its requests, stall_no_inst 10 %):
// the compute role's straight-line code is ~57 KB per item, more than the SM-level instruction cache holds, and the two
// co-resident CTAs walk it out of phase.  The looped form is ~4x smaller.
#ifndef C2_LOOP
#define C2_LOOP 0
#endif
static_assert(!C2_LOOP || (C2_FFMA2 && !C2_SPLIT_BAR && !C2_SKIP_OUTER), "looped hops: packed arithmetic, one barrier per hop");
static_assert(!C2_IO_OUT || (C2_FFMA2 && C2_CPT == 4 && C2_BC == 3 && !C2_IO_DRAIN && !C2_SPLIT_BAR && !C2_LOOP),
              "IO-side basis stores: implemented for the default block shape and hop sequence");
#if C2_SYMW
// the weight of the tap (r, cc) <- (sr, sc), both inside the block, is read from the OTHER pixel's table entry
__host__ __device__ constexpr bool sym_other(int r, int cc, int sr, int sc) { return (sr * C2_BC + sc) < (r * C2_BC + cc); }
#endif

// One hop on the thread's 3x3 block, in two parts so that the part that needs no other thread's data overlaps the
// wait for the neighbours:
//   hop_inside:    acc <- w_c * in - (HAS_OLD ? acc : 0) + taps whose source pixel lies inside the block (49 of 81)
//   hop_perimeter: acc += taps whose source is one of the 16 perimeter pixels (loaded from `src`, which points at the
//                  thread's own (r = 0, cc = 0) position of the buffer holding `in` of all threads); halved if HALVE
template <bool HAS_OLD, int ROT>
__device__ __forceinline__ void hop_inside(const cvec (&in)[C2_BR][C2_BC], cvec (&acc)[C2_BR][C2_BC],
                                           const float (&w)[C2_BR][C2_BC][9], const float wdiag) {
#pragma unroll
  for (int r = 0; r < C2_BR; ++r)
#pragma unroll
    for (int cc = 0; cc < C2_BC; ++cc) {
      const float wc = C2_CDIAG ? wdiag : w[r][cc][8];
      if (HAS_OLD) f4_fms<ROT>(wc, in[r][cc], acc[r][cc]);
      else f4_mul<ROT>(wc, in[r][cc], acc[r][cc]);
    }
#pragma unroll
  for (int dr = -1; dr <= 1; ++dr)
#pragma unroll
    for (int dc = -1; dc <= 1; ++dc)
#pragma unroll
      for (int r = 0; r < C2_BR; ++r)
#pragma unroll
        for (int cc = 0; cc < C2_BC; ++cc) {
          if (dr == 0 && dc == 0) continue;
          const int sr = r + dr, sc = cc + dc;
          if (!(sr >= 0 && sr < C2_BR && sc >= 0 && sc < C2_BC)) continue;
#if C2_SYMW
          f4_fma<ROT>(sym_other(r, cc, sr, sc) ? w[sr][sc][dir_of(-dr, -dc)] : w[r][cc][dir_of(dr, dc)], in[sr][sc],
                      acc[r][cc]);
#else
          f4_fma<ROT>(w[r][cc][dir_of(dr, dc)], in[sr][sc], acc[r][cc]);
#endif
        }
}
template <bool HALVE, int ROT>
__device__ __forceinline__ void hop_perimeter(cvec (&acc)[C2_BR][C2_BC], const float (&w)[C2_BR][C2_BC][9],
                                              const cvec* __restrict__ src) {
  // position offsets of columns -1, 0, .., C2_BC relative to the own column-0 position (k = column + 1)
  auto co = [](int k) constexpr { return k == 0 ? (C2_BC - 1) * C2_NB - 1 : (k == C2_BC + 1 ? 1 : (k - 1) * C2_NB); };
  cvec top[C2_BC + 2], bot[C2_BC + 2], lft[C2_BR], rgt[C2_BR];
#pragma unroll
  for (int k = 0; k < C2_BC + 2; ++k) top[k] = src[(-C2_LW + co(k)) * C2_VP];
#pragma unroll
  for (int r = 0; r < C2_BR; ++r) {
    lft[r] = src[(r * C2_LW + co(0)) * C2_VP];
    rgt[r] = src[(r * C2_LW + 1) * C2_VP];
  }
#pragma unroll
  for (int k = 0; k < C2_BC + 2; ++k) bot[k] = src[(C2_BR * C2_LW + co(k)) * C2_VP];
#pragma unroll
  for (int dr = -1; dr <= 1; ++dr)
#pragma unroll
    for (int dc = -1; dc <= 1; ++dc)
#pragma unroll
      for (int r = 0; r < C2_BR; ++r)
#pragma unroll
        for (int cc = 0; cc < C2_BC; ++cc) {
          if (dr == 0 && dc == 0) continue;
          const int sr = r + dr, sc = cc + dc;
          if (sr >= 0 && sr < C2_BR && sc >= 0 && sc < C2_BC) continue;
          const float wv = w[r][cc][dir_of(dr, dc)];
          if (sr < 0) f4_fma<ROT>(wv, top[sc + 1], acc[r][cc]);
          else if (sr > C2_BR - 1) f4_fma<ROT>(wv, bot[sc + 1], acc[r][cc]);
          else if (sc < 0) f4_fma<ROT>(wv, lft[sr], acc[r][cc]);
          else f4_fma<ROT>(wv, rgt[sr], acc[r][cc]);
        }
  if (HALVE) {
#pragma unroll
    for (int r = 0; r < C2_BR; ++r)
#pragma unroll
      for (int cc = 0; cc < C2_BC; ++cc) acc[r][cc] = f4_scale(0.5f, acc[r][cc]);
  }
}

// TMEM -> bias / activation -> y for the 32 accumulator rows (TMEM lane quarter `warp & 3`) of the 3 M-tiles; rows[mt] = row
// of y of this thread's accumulator row (or -1).  Warp-collective.
__device__ __forceinline__ void drain_quarter(const Conv2Args& a, uint32_t tmem_base, int warp, const int (&rows)[3], int64_t b) {
  const int N = a.N, NV16 = N / 16;
#pragma unroll 1
  for (int mt = 0; mt < 3; ++mt) {
    const int row = rows[mt];
    float* yrow = a.y + (b * a.M + (row >= 0 ? row : 0)) * (int64_t)N;
#pragma unroll 1
    for (int cc = 0; cc < NV16; ++cc) {
      uint32_t r[16];
      ptx::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(mt * N + cc * 16), r);
      ptx::tmem_ld_wait();
      if (row >= 0) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          float4 o = make_float4(__uint_as_float(r[v * 4]), __uint_as_float(r[v * 4 + 1]), __uint_as_float(r[v * 4 + 2]),
                                 __uint_as_float(r[v * 4 + 3]));
          if (a.bias != nullptr) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(a.bias + cc * 16 + v * 4));
            o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
          }
          if (a.act != DS_ACT_LINEAR) o = act4(o, a.act);
          __stcs(reinterpret_cast<float4*>(yrow + cc * 16 + v * 4), o);
        }
      }
    }
  }
}

template <bool CHEB>
__global__ void __launch_bounds__(C2_THREADS, 2) lattice_conv2_kernel(const Conv2Args a) {
  C2_DYNAMIC_SMEM(c2_smem);
  float4* const bufs = reinterpret_cast<float4*>(c2_smem);  // S0, S1 (input staging), X0, X1 (hop results)
  const int N = a.N, nsteps = a.nsteps;
  const uint32_t img_bytes = (uint32_t)N * C2_FC * 4;                 // one (chunk, hop) weight image
  const uint32_t wslice_bytes = (uint32_t)(nsteps + 1) * img_bytes;  // all hops of one chunk
  uint8_t* const wbuf = c2_smem + (size_t)4 * C2_BUF * 16;
  int32_t* const s_pix = reinterpret_cast<int32_t*>(wbuf + 2 * (size_t)wslice_bytes);
  Conv2Ctl* const ctl = reinterpret_cast<Conv2Ctl*>(s_pix + C2_P);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_chunks = a.F / C2_FC;
  const int64_t b_per = (a.B + a.b_split - 1) / a.b_split;
  const int n_units = a.n_tiles * a.b_split;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&ctl->in_full[i], C2_NLOAD);
      ptx::mbar_init(&ctl->in_empty[i], 1);
      ptx::mbar_init(&ctl->w_full[i], 1);
      ptx::mbar_init(&ctl->item_done[i], 1);
      ptx::mbar_init(&ctl->hop_full[i], C2_NCOMP);
#if C2_SPLIT_BAR
      ptx::mbar_init(&ctl->hop_ready[i], C2_NCOMP);
#endif
      ptx::mbar_init(&ctl->mma_done[i], 1);
#if C2_IO_OUT
      ptx::mbar_init(&ctl->out_done[i], C2_NLOAD);
#endif
    }
    ptx::mbar_init(&ctl->acc_full, 1);
    ptx::mbar_init(&ctl->acc_empty, C2_IO_DRAIN ? 4 : C2_NEPI);
    ptx::fence_mbar_init();
  }
  if (warp == C2_NCW) ptx::tmem_alloc(&ctl->tmem_base, C2_TMEM_COLS);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp < C2_NCW) {
    // ================================ compute warps ================================
    C2_SETMAXNREG_INC(C2_REG_COMPUTE);
    // (skipping the block rows that lie outside the valid region of hops 3, 4 was measured: no gain, the kernel is
    // bound by the per-hop dependency chain, not by issue slots)
#if C2_CPT == 2
    // lane = (quad q, column block cb, half h): the 16 lanes of a half-warp touch 128 consecutive bytes (LDS.64 / STS.64)
#if C2_BC == 6
    // lane = (block-row half, quad q, column block cb (4), half h); warp 0 owns the outermost block rows 0 and 7
    const int R = warp == 0 ? 7 * (lane >> 4) : 2 * warp - 1 + (lane >> 4), q = (lane >> 3) & 1, cb = (lane >> 1) & 3, h = lane & 1;
#else
    const int R = warp, q = lane >> 4, cb = (lane >> 1) & 7, h = lane & 1;
#endif
#else
    const int R = 2 * warp + (lane >> 4), q = (lane >> 3) & 1, cb = lane & 7, h = 0;
#endif
    const int own0 = (q * C2_PL + (C2_BR * R + 1) * C2_LW + cb) * C2_VP + h;  // cvec index of the own (0, 0) position
    const int FV = a.F / C2_CPT, NV16 = N / 16;
    const int egrp = C2_NEPI == 8 ? (warp >> 2) : 0;  // which half of the accumulator slices this warp drains
    const bool has_out = a.out[0] != nullptr || a.out[1] != nullptr || a.out[2] != nullptr || a.out[3] != nullptr;
    float w[C2_BR][C2_BC][9];
    cvec A[C2_BR][C2_BC], Bv[C2_BR][C2_BC];
    uint32_t it = 0, g = 0;
    uint32_t cnt_done[2] = {0, 0}, par[2] = {0, 0};
    int last_bar = -1;
    uint32_t last_par = 0;
    int erow[3] = {-1, -1, -1};
    bool pend = false;
    uint32_t pend_g = 0;
    int64_t pend_b = 0;
    int pend_rows[3] = {-1, -1, -1};

    // drain the accumulators of group pend_g: TMEM -> bias/activation -> y (own pixels of the tile)
    auto epilogue = [&]() {
      ptx::mbar_wait(&ctl->acc_full, pend_g & 1);
      ptx::tc_fence_after_sync();
#if C2_EPI_PIPE
      // one flat loop over this warp's share of the 3 * N / 16 accumulator slices (slice = egrp + j * groups) with the
      // tensor-memory load of the next slice in flight while the current one is biased, activated and stored (the ncu
      // source page shows the drain waiting on every tcgen05.ld)
      {
        constexpr int G = C2_NEPI / 4;
        const int n_slices = 3 * NV16;
        const uint32_t t0 = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t rbuf[2][16];
        if (egrp < n_slices) {
          ptx::tmem_ld_32x32b_x16(t0 + (uint32_t)((egrp / NV16) * N + (egrp % NV16) * 16), rbuf[0]);
          ptx::tmem_ld_wait();
        }
#pragma unroll 1
        for (int i = egrp; i < n_slices; i += 2 * G) {
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const int cur = i + h2 * G;
            if (cur >= n_slices) break;
            const int mt = cur / NV16, cc = cur - mt * NV16;
            if (cur + G < n_slices) {
              const int nmt = (cur + G) / NV16, ncc = (cur + G) - nmt * NV16;
              ptx::tmem_ld_32x32b_x16(t0 + (uint32_t)(nmt * N + ncc * 16), rbuf[h2 ^ 1]);
            }
            const int row = pend_rows[mt];
            if (row >= 0) {
              float* yrow = a.y + (pend_b * a.M + row) * (int64_t)N;
#pragma unroll
              for (int v = 0; v < 4; ++v) {
                float4 o = make_float4(__uint_as_float(rbuf[h2][v * 4]), __uint_as_float(rbuf[h2][v * 4 + 1]),
                                       __uint_as_float(rbuf[h2][v * 4 + 2]), __uint_as_float(rbuf[h2][v * 4 + 3]));
                if (a.bias != nullptr) {
                  const float4 bv = __ldg(reinterpret_cast<const float4*>(a.bias + cc * 16 + v * 4));
                  o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
                }
                if (a.act != DS_ACT_LINEAR) o = act4(o, a.act);
                __stcs(reinterpret_cast<float4*>(yrow + cc * 16 + v * 4), o);
              }
            }
            ptx::tmem_ld_wait();
          }
        }
      }
#else
#pragma unroll 1
      for (int mt = 0; mt < 3; ++mt) {
        const int row = pend_rows[mt];
        float* yrow = a.y + (pend_b * a.M + (row >= 0 ? row : 0)) * (int64_t)N;
#pragma unroll 1
        for (int cc = 0; cc < NV16; ++cc) {
          if (C2_NEPI == 8 && ((mt * NV16 + cc) & 1) != egrp) continue;
          uint32_t r[16];
          ptx::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(mt * N + cc * 16), r);
          ptx::tmem_ld_wait();
          if (row >= 0) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              float4 o = make_float4(__uint_as_float(r[v * 4]), __uint_as_float(r[v * 4 + 1]), __uint_as_float(r[v * 4 + 2]),
                                     __uint_as_float(r[v * 4 + 3]));
              if (a.bias != nullptr) {
                const float4 bv = __ldg(reinterpret_cast<const float4*>(a.bias + cc * 16 + v * 4));
                o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
              }
              if (a.act != DS_ACT_LINEAR) o = act4(o, a.act);
              __stcs(reinterpret_cast<float4*>(yrow + cc * 16 + v * 4), o);
            }
          }
        }
      }
#endif
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&ctl->acc_empty);
      pend = false;
    };

    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      const int tile = unit / a.b_split;
      const int64_t b_begin = (int64_t)(unit % a.b_split) * b_per;
      const int64_t b_end = min(a.B, b_begin + b_per);
      if (b_begin >= b_end) continue;
      const int32_t* tpix = a.pix + (size_t)tile * C2_P;
      if (has_out && !C2_IO_OUT) {  // own-pixel rows of this tile for the basis stores
        ptx::named_bar_sync(1, C2_NCOMP);
        for (int p = tid; p < C2_P; p += C2_NCOMP) {
          const int j = p / C2_LW, c = p % C2_LW;
          s_pix[p] = (j >= C2_H && j < C2_H + C2_T && c >= C2_H && c < C2_H + C2_T) ? __ldg(tpix + p) : -1;
        }
        ptx::named_bar_sync(1, C2_NCOMP);
      }
#pragma unroll
      for (int r = 0; r < C2_BR; ++r)
#pragma unroll
        for (int cc = 0; cc < C2_BC; ++cc) {
          const float* wp = a.w + ((size_t)tile * C2_P + (size_t)(C2_BR * R + r) * C2_LW + C2_BC * cb + cc) * 9;
#if C2_SYMW
#pragma unroll
          for (int dr = -1; dr <= 1; ++dr)
#pragma unroll
            for (int dc = -1; dc <= 1; ++dc) {
              const int sr = r + dr, sc = cc + dc, d = dir_of(dr, dc);
              const bool inside = sr >= 0 && sr < C2_BR && sc >= 0 && sc < C2_BC && d != 8;
              if (inside && sym_other(r, cc, sr, sc)) continue;  // kept by the other pixel of the link
              if (C2_CDIAG && d == 8) continue;
              C2_SCALED_WEIGHT(w[r][cc][d], __ldg(wp + d), a.wscale);
            }
#else
#pragma unroll
          for (int d = 0; d < (C2_CDIAG ? 8 : 9); ++d) C2_SCALED_WEIGHT(w[r][cc][d], __ldg(wp + d), a.wscale);
#endif
        }
#pragma unroll
      for (int mt = 0; mt < 3; ++mt) {  // accumulator row (mt, TMEM lane) -> lattice position -> row of y
        const int m = mt * 128 + (warp & 3) * 32 + lane;
        const int j = C2_H + m / C2_LW, p = m % C2_LW, c = col_of_slot(p);
        erow[mt] = (c >= C2_H && c < C2_H + C2_T) ? __ldg(tpix + j * C2_LW + c) : -1;
      }

      for (int64_t b = b_begin; b < b_end; ++b) {
        for (int c = 0; c < n_chunks; ++c) {
          const uint32_t st = it & 1;
          const cvec* S = reinterpret_cast<const cvec*>(bufs + (size_t)st * C2_BUF);
          cvec* X[2] = {reinterpret_cast<cvec*>(bufs + 2 * (size_t)C2_BUF), reinterpret_cast<cvec*>(bufs + 3 * (size_t)C2_BUF)};
          if (C2_PROBE && a.dbg != nullptr && blockIdx.x == 0 && it < 16 && tid == 0) a.dbg[((size_t)(it & 15) * 5) * 8 + 0] = clock64();
          ptx::mbar_wait(&ctl->in_full[st], (it >> 1) & 1);
          if (C2_PROBE && a.dbg != nullptr && blockIdx.x == 0 && it < 16 && tid == 0) a.dbg[((size_t)(it & 15) * 5) * 8 + 1] = clock64();
#pragma unroll
          for (int r = 0; r < C2_BR; ++r)
#pragma unroll
            for (int cc = 0; cc < C2_BC; ++cc) A[r][cc] = S[own0 + (r * C2_LW + cc * C2_NB) * C2_VP];

#if C2_LOOP
          // Hop s: `in` (T_{s-1}) -> `acc` (T_{s-2} -> T_s); one body per buffer parity p = (s - 1) & 1 (the register
          // arrays alternate), everything that depends on s is a run-time value.  Chebyshev: `acc` starts as zero, so
          // hop 1 is the same "w * in - acc" as the others (T_1 = (2 L~ T_0) / 2: halved after the perimeter taps).
          auto hop_body = [&](auto p_tag, const cvec(&in)[C2_BR][C2_BC], cvec(&acc)[C2_BR][C2_BC], const int s) {
            constexpr int p = decltype(p_tag)::value;
            hop_inside<CHEB, 2>(in, acc, w, a.wdiag);  // own registers only: runs before the wait for the neighbours
            const cvec* src = S;
            if (s > 1) {
              ptx::mbar_wait(&ctl->hop_full[p ^ 1], par[p ^ 1]);  // hop s - 1 published by everybody
              src = X[p ^ 1];
            }
            // the UMMAs that read X[p] two hops ago: probe now, consume after the arithmetic (hides the round trip)
            const bool mma_ok = cnt_done[p] == 0 || ptx::mbar_test_wait(&ctl->mma_done[p], (cnt_done[p] - 1) & 1);
            hop_perimeter<false, 2>(acc, w, src + own0);
            if (CHEB && s == 1) {
#pragma unroll
              for (int r = 0; r < C2_BR; ++r)
#pragma unroll
                for (int cc = 0; cc < C2_BC; ++cc) acc[r][cc] = f4_scale(0.5f, acc[r][cc]);
            }
            if (!mma_ok) ptx::mbar_wait(&ctl->mma_done[p], (cnt_done[p] - 1) & 1);
            if (s == 1 && last_bar >= 0) {  // previous item's last phase: everybody has finished reading X[0]
              ptx::mbar_wait(&ctl->hop_full[last_bar], last_par);
              last_bar = -1;
            }
            cvec* dst = X[p] + own0;
#pragma unroll
            for (int r = 0; r < C2_BR; ++r)
#pragma unroll
              for (int cc = 0; cc < C2_BC; ++cc) dst[(r * C2_LW + cc * C2_NB) * C2_VP] = acc[r][cc];
            ptx::fence_proxy_async_smem();
            ptx::mbar_arrive(&ctl->hop_full[p]);
            if (has_out) {
              float* outp = s == 1 ? a.out[0] : (s == 2 ? a.out[1] : (s == 3 ? a.out[2] : a.out[3]));
              if (outp != nullptr) {
                cvec* ob = reinterpret_cast<cvec*>(outp + (b * a.M * a.F + c * C2_FC)) + q * C2_VP + h;
#pragma unroll
                for (int r = 0; r < C2_BR; ++r)
#pragma unroll
                  for (int cc = 0; cc < C2_BC; ++cc) {
                    const int row = s_pix[(C2_BR * R + r) * C2_LW + C2_BC * cb + cc];
                    if (row >= 0) __stcs(ob + (int64_t)row * FV, acc[r][cc]);
                  }
              }
            }
            par[p] = cnt_done[p] & 1;  // parity of the phase this arrival belongs to
            cnt_done[p]++;
          };
          if (CHEB) {
#pragma unroll
            for (int r = 0; r < C2_BR; ++r)
#pragma unroll
              for (int cc = 0; cc < C2_BC; ++cc) Bv[r][cc] = cvec{};
          }
          int last = 0;
#pragma unroll 1
          for (int s = 1; s <= nsteps; s += 2) {
            hop_body(std::integral_constant<int, 0>{}, A, Bv, s);
            last = 0;
            if (s == 1 && pend) epilogue();
            if (s + 1 <= nsteps) {
              hop_body(std::integral_constant<int, 1>{}, Bv, A, s + 1);
              last = 1;
            }
          }
#else
          // Hop s: `in` (T_{s-1}) -> `acc` (T_{s-2} -> T_s), the register arrays alternate; results are published in
          // X[(s-1)&1].  There is no __syncthreads: every thread arrives on hop_full[p] (count 128) after its stores
          // and proxy fence, starts the inside part of the NEXT hop (own registers only, 60 % of the taps) and only
          // then waits for the phase before it loads its perimeter.  Every thread computes its whole block on every
          // hop: positions outside the shrinking valid region hold don't-care values that never reach a valid
          // output (a valid output only reads valid inputs).
          // (st.async + complete_tx instead of STS + proxy fence was measured: ~20 B/clk, 3x slower.)
          // C2_SKIP_OUTER: hop s of an item with `nsteps` hops needs lattice rows 4 - (nsteps - s) .. 19 + (nsteps - s)
          auto skip_hop = [&](int s) { return C2_SKIP_OUTER && C2_BC == 6 && warp == 0 && s > nsteps - 2; };
          auto finish_hop = [&](auto s_tag, cvec(&acc)[C2_BR][C2_BC], const cvec* src) {
            constexpr int s = decltype(s_tag)::value;
            constexpr int p = (s - 1) & 1;
            const bool probe = C2_PROBE && a.dbg != nullptr && blockIdx.x == 0 && it < 16 && tid == 0;
            long long* pd = a.dbg + ((size_t)(it & 15) * 5 + s) * 8;
            if (probe) pd[0] = clock64();
            // the UMMAs that read X[p] two hops ago: probe now, consume after the arithmetic (hides the round trip)
            const bool skp = skip_hop(s);  // warp-uniform
            const bool mma_ok = skp || cnt_done[p] == 0 || ptx::mbar_test_wait(&ctl->mma_done[p], (cnt_done[p] - 1) & 1);
            if (!skp) hop_perimeter<(CHEB && s == 1), rot_of_hop(s)>(acc, w, src + own0);
            if (probe) pd[1] = clock64();
            if (!mma_ok) ptx::mbar_wait(&ctl->mma_done[p], (cnt_done[p] - 1) & 1);
#if C2_IO_OUT
            // ... and the gather warps that copied X[p]'s previous contents to the basis tensor
            if (has_out && cnt_done[p] > 0) ptx::mbar_wait(&ctl->out_done[p], (cnt_done[p] - 1) & 1);
#endif
            if (s == 1 && last_bar >= 0) {  // previous item's last phase: everybody has finished reading X[0] (a warp that
                                            // sits this hop out waits too: its arrival must not land in that phase)
#if C2_SPLIT_BAR
              ptx::mbar_wait(&ctl->hop_ready[last_bar], last_par);
#else
              ptx::mbar_wait(&ctl->hop_full[last_bar], last_par);
#endif
              last_bar = -1;
            }
            if (probe) pd[2] = clock64();
            cvec* dst = X[p] + own0;
            if (!skp) {
#pragma unroll
              for (int r = 0; r < C2_BR; ++r)
#pragma unroll
                for (int cc = 0; cc < C2_BC; ++cc) dst[(r * C2_LW + cc * C2_NB) * C2_VP] = acc[r][cc];
            }
#if C2_SPLIT_BAR
            ptx::mbar_arrive(&ctl->hop_ready[p]);  // fence + hop_full arrive: publish(p), after the next inside taps
#else
#if !C2_FENCE_BY_ISSUER
            ptx::fence_proxy_async_smem();
#endif
            ptx::mbar_arrive(&ctl->hop_full[p]);
#endif
            if (probe) pd[3] = clock64();
            float* outp = C2_IO_OUT ? nullptr : a.out[s - 1];
            if (outp != nullptr && !skp) {
              cvec* ob = reinterpret_cast<cvec*>(outp + (b * a.M * a.F + c * C2_FC)) + q * C2_VP + h;
#pragma unroll
              for (int r = 0; r < C2_BR; ++r)
#pragma unroll
                for (int cc = 0; cc < C2_BC; ++cc) {
                  const int row = s_pix[(C2_BR * R + r) * C2_LW + C2_BC * cb + cc];
#if C2_CPT == 4
                  const float4 v = hop_is_rotated(s) ? make_float4(acc[r][cc].w, acc[r][cc].x, acc[r][cc].y, acc[r][cc].z) : acc[r][cc];
#else
                  const cvec v = acc[r][cc];
#endif
                  if (row >= 0) __stcs(ob + (int64_t)row * FV, v);
                }
            }
            par[p] = cnt_done[p] & 1;  // parity of the phase this arrival belongs to
            cnt_done[p]++;
            if (probe) pd[4] = clock64();
          };
          // wait until hop s has been published by everybody (its phase on hop_full[(s-1)&1])
#if C2_SPLIT_BAR
          auto wait_hop = [&](int pp) { ptx::mbar_wait(&ctl->hop_ready[pp], par[pp]); };
          // hand the hop in X[pp] to the UMMA issuer: this thread's stores -> async proxy, then hop_full
          auto publish = [&](int pp) {
#if !C2_FENCE_BY_ISSUER
            ptx::fence_proxy_async_smem();
#endif
            ptx::mbar_arrive(&ctl->hop_full[pp]);
          };
          if (!skip_hop(1)) hop_inside<false, rot_of_hop(1)>(A, Bv, w, a.wdiag);
          finish_hop(std::integral_constant<int, 1>{}, Bv, S);
          bool pub = false;
          if (pend) {  // the drain is long: do not hold hop 1 back behind it
            publish(0);
            pub = true;
            epilogue();
          }
          int last = 0;
          if (nsteps >= 2) {
            if (!skip_hop(2)) hop_inside<CHEB, rot_of_hop(2)>(Bv, A, w, a.wdiag);
            if (!pub) publish(0);
            wait_hop(0);
            finish_hop(std::integral_constant<int, 2>{}, A, X[0]);
            last = 1;
            if (nsteps == 2) publish(1);
          } else if (!pub) {
            publish(0);
          }
          if (nsteps >= 3) {
            if (!skip_hop(3)) hop_inside<CHEB, rot_of_hop(3)>(A, Bv, w, a.wdiag);
            publish(1);
            wait_hop(1);
            finish_hop(std::integral_constant<int, 3>{}, Bv, X[1]);
            last = 0;
            if (nsteps == 3) publish(0);
          }
          if (nsteps >= 4) {
            if (!skip_hop(4)) hop_inside<CHEB, rot_of_hop(4)>(Bv, A, w, a.wdiag);
            publish(0);
            wait_hop(0);
            finish_hop(std::integral_constant<int, 4>{}, A, X[0]);
            last = 1;
            publish(1);
          }
#else
          auto wait_hop = [&](int pp) { ptx::mbar_wait(&ctl->hop_full[pp], par[pp]); };

          if (!skip_hop(1)) hop_inside<false, rot_of_hop(1)>(A, Bv, w, a.wdiag);
          finish_hop(std::integral_constant<int, 1>{}, Bv, S);
          if (pend) epilogue();
          int last = 0;
          if (nsteps >= 2) {
            if (!skip_hop(2)) hop_inside<CHEB, rot_of_hop(2)>(Bv, A, w, a.wdiag);
            wait_hop(0);
            finish_hop(std::integral_constant<int, 2>{}, A, X[0]);
            last = 1;
          }
          if (nsteps >= 3) {
            if (!skip_hop(3)) hop_inside<CHEB, rot_of_hop(3)>(A, Bv, w, a.wdiag);
            wait_hop(1);
            finish_hop(std::integral_constant<int, 3>{}, Bv, X[1]);
            last = 0;
          }
          if (nsteps >= 4) {
            if (!skip_hop(4)) hop_inside<CHEB, rot_of_hop(4)>(Bv, A, w, a.wdiag);
            wait_hop(0);
            finish_hop(std::integral_constant<int, 4>{}, A, X[0]);
            last = 1;
          }
#endif
#endif  // C2_LOOP
          last_bar = last;
          last_par = par[last];

          if (c == n_chunks - 1) {
            pend = !C2_IO_DRAIN && warp < C2_NEPI;  // TMEM lanes 32 (w % 4) .. + 31 belong to warp w: 4 (or 2 x 4, half the slices each) warps drain
            pend_g = g;
            pend_b = b;
            pend_rows[0] = erow[0]; pend_rows[1] = erow[1]; pend_rows[2] = erow[2];
            ++g;
          }
          ++it;
        }
      }
    }
    if (pend) epilogue();
  } else if (warp == C2_NCW) {
#if C2_IO_DRAIN
    // ================================ UMMA issuer / weight streamer (+ drain of TMEM lane quarter 0) =================
    C2_SETMAXNREG_DEC(C2_REG_IO);
    {
      // Lane 0 alone waits, issues and commits (a lane that polled the pipeline barriers without gating them could miss
      // a phase); the other lanes follow the item sequence without touching a barrier and meet lane 0 at the end of every
      // group, where the whole warp drains quarter 0: acc_full cannot advance again before this warp has arrived on
      // acc_empty, so that wait is safe for all lanes.
      uint32_t total_items = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int64_t b_begin = (int64_t)(unit % a.b_split) * b_per;
        const int64_t b_end = min(a.B, b_begin + b_per);
        if (b_end > b_begin) total_items += (uint32_t)((b_end - b_begin) * n_chunks);
      }
      const uint32_t idesc = ptx::make_idesc_tf32(128, N, 0, 0);
      const uint32_t buf_u32 = ptx::smem_u32(bufs);
      const uint32_t wbuf_u32 = ptx::smem_u32(wbuf);
      auto load_w = [&](uint32_t item, int chunk) {
        const uint32_t wb = item & 1;
        ptx::mbar_arrive_expect_tx(&ctl->w_full[wb], wslice_bytes);
        ptx::bulk_load_1d(wbuf + (size_t)wb * wslice_bytes,
                          reinterpret_cast<const uint8_t*>(a.b_img) + (size_t)chunk * wslice_bytes, wslice_bytes,
                          &ctl->w_full[wb]);
      };
      auto issue = [&](uint32_t a_buf_u32, uint32_t w_u32, int k, bool first) {
        const uint32_t a_base = a_buf_u32 + (uint32_t)((C2_H + 1) * C2_LW) * 16;  // plane 0, lattice row 4, position 0
        const uint64_t bd = ptx::make_smem_desc(w_u32 + (uint32_t)k * img_bytes, (uint32_t)N * 16, 128,
                                                ptx::LAYOUT_SWIZZLE_NONE);
#pragma unroll
        for (int mt = 0; mt < 3; ++mt) {
          const uint64_t ad = ptx::make_smem_desc(a_base + (uint32_t)(mt * 128) * 16, C2_PL * 16, 128,
                                                  ptx::LAYOUT_SWIZZLE_NONE);
          ptx::umma_tf32(tmem_base + (uint32_t)(mt * N), ad, bd, idesc, first ? 0u : 1u);
        }
      };
      uint32_t it = 0, g = 0, ch[2] = {0, 0};
      if (lane == 0 && total_items > 0) load_w(0, 0);
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int64_t b_begin = (int64_t)(unit % a.b_split) * b_per;
        const int64_t b_end = min(a.B, b_begin + b_per);
        if (b_begin >= b_end) continue;
        int erow[3];
#pragma unroll
        for (int mt = 0; mt < 3; ++mt) {  // accumulator row (mt, TMEM lane of quarter 0) -> lattice position -> row of y
          const int m = mt * 128 + lane;
          const int j = C2_H + m / C2_LW, p = m % C2_LW, cl = col_of_slot(p);
          erow[mt] = (cl >= C2_H && cl < C2_H + C2_T) ? __ldg(a.pix + (size_t)(unit / a.b_split) * C2_P + j * C2_LW + cl) : -1;
        }
        for (int64_t b = b_begin; b < b_end; ++b) {
          for (int c = 0; c < n_chunks; ++c) {
            if (lane == 0) {
              const uint32_t st = it & 1;
              if (c == 0 && g >= 1) ptx::mbar_wait_backoff(&ctl->acc_empty, (g - 1) & 1, (uint32_t)a.sleep_mma);  // previous group drained
              ptx::mbar_wait_backoff(&ctl->w_full[st], (it >> 1) & 1, (uint32_t)a.sleep_mma);
              ptx::mbar_wait_backoff(&ctl->in_full[st], (it >> 1) & 1, (uint32_t)a.sleep_mma);
              ptx::tc_fence_after_sync();
              const uint32_t w_u32 = wbuf_u32 + st * wslice_bytes;
              issue(buf_u32 + st * (uint32_t)(C2_BUF * 16), w_u32, 0, c == 0);
              if (it + 1 < total_items) {  // stream the next chunk's weight slice into the other buffer
                if (it >= 1) ptx::mbar_wait_backoff(&ctl->item_done[(it + 1) & 1], ((it - 1) >> 1) & 1, (uint32_t)a.sleep_mma);
                load_w(it + 1, (c + 1) % n_chunks);
              }
              for (int s = 1; s <= nsteps; ++s) {
                const int p = (s - 1) & 1;
                ptx::mbar_wait_backoff(&ctl->hop_full[p], ch[p] & 1, (uint32_t)a.sleep_mma);
                ch[p]++;
                ptx::tc_fence_after_sync();
                issue(buf_u32 + (uint32_t)(2 + p) * (uint32_t)(C2_BUF * 16), w_u32, s, false);
                ptx::umma_commit(&ctl->mma_done[p]);
                if (s == 1) ptx::umma_commit(&ctl->in_empty[st]);
              }
              ptx::umma_commit(&ctl->item_done[st]);
              if (c == n_chunks - 1) ptx::umma_commit(&ctl->acc_full);
              ++it;
            }
            if (c == n_chunks - 1) {
              __syncwarp();
              ptx::mbar_wait_backoff(&ctl->acc_full, g & 1, (uint32_t)a.sleep_mma);
              ptx::tc_fence_after_sync();
              drain_quarter(a, tmem_base, warp, erow, b);
              ptx::tc_fence_before_sync();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(&ctl->acc_empty);
              ++g;
            }
          }
        }
      }
    }
    __syncwarp();
#else
    // ================================ UMMA issuer / weight streamer ================================
    C2_SETMAXNREG_DEC(C2_REG_IO);
    if (lane == 0) {
      uint32_t total_items = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int64_t b_begin = (int64_t)(unit % a.b_split) * b_per;
        const int64_t b_end = min(a.B, b_begin + b_per);
        if (b_end > b_begin) total_items += (uint32_t)((b_end - b_begin) * n_chunks);
      }
      const uint32_t idesc = ptx::make_idesc_tf32(128, N, 0, 0);
      const uint32_t buf_u32 = ptx::smem_u32(bufs);
      const uint32_t wbuf_u32 = ptx::smem_u32(wbuf);
      auto load_w = [&](uint32_t item, int chunk) {
        const uint32_t wb = item & 1;
        ptx::mbar_arrive_expect_tx(&ctl->w_full[wb], wslice_bytes);
        ptx::bulk_load_1d(wbuf + (size_t)wb * wslice_bytes,
                          reinterpret_cast<const uint8_t*>(a.b_img) + (size_t)chunk * wslice_bytes, wslice_bytes,
                          &ctl->w_full[wb]);
      };
      // A = positions of lattice rows 4..19 (3 M-tiles of 128) x the item's 8 channels (one K step = 2 planes)
      auto issue = [&](uint32_t a_buf_u32, uint32_t w_u32, int k, bool first) {
        const uint32_t a_base = a_buf_u32 + (uint32_t)((C2_H + 1) * C2_LW) * 16;  // plane 0, lattice row 4, position 0
        const uint64_t bd = ptx::make_smem_desc(w_u32 + (uint32_t)k * img_bytes, (uint32_t)N * 16, 128,
                                                ptx::LAYOUT_SWIZZLE_NONE);
#pragma unroll
        for (int mt = 0; mt < 3; ++mt) {
          const uint64_t ad = ptx::make_smem_desc(a_base + (uint32_t)(mt * 128) * 16, C2_PL * 16, 128,
                                                  ptx::LAYOUT_SWIZZLE_NONE);
          ptx::umma_tf32(tmem_base + (uint32_t)(mt * N), ad, bd, idesc, first ? 0u : 1u);
        }
      };
      uint32_t it = 0, g = 0, ch[2] = {0, 0};
      if (total_items > 0) load_w(0, 0);
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int64_t b_begin = (int64_t)(unit % a.b_split) * b_per;
        const int64_t b_end = min(a.B, b_begin + b_per);
        if (b_begin >= b_end) continue;
        for (int64_t b = b_begin; b < b_end; ++b) {
          for (int c = 0; c < n_chunks; ++c) {
            const uint32_t st = it & 1;
            if (c == 0 && g >= 1) ptx::mbar_wait_backoff(&ctl->acc_empty, (g - 1) & 1, (uint32_t)a.sleep_mma);  // previous group drained
            ptx::mbar_wait_backoff(&ctl->w_full[st], (it >> 1) & 1, (uint32_t)a.sleep_mma);
            ptx::mbar_wait_backoff(&ctl->in_full[st], (it >> 1) & 1, (uint32_t)a.sleep_mma);
            ptx::tc_fence_after_sync();
            const uint32_t w_u32 = wbuf_u32 + st * wslice_bytes;
            issue(buf_u32 + st * (uint32_t)(C2_BUF * 16), w_u32, 0, c == 0);
            if (it + 1 < total_items) {  // stream the next chunk's weight slice into the other buffer
              if (it >= 1) ptx::mbar_wait_backoff(&ctl->item_done[(it + 1) & 1], ((it - 1) >> 1) & 1, (uint32_t)a.sleep_mma);
              load_w(it + 1, (c + 1) % n_chunks);
            }
            for (int s = 1; s <= nsteps; ++s) {
              const int p = (s - 1) & 1;
              ptx::mbar_wait_backoff(&ctl->hop_full[p], ch[p] & 1, (uint32_t)a.sleep_mma);
              const bool probe = C2_PROBE && a.dbg != nullptr && blockIdx.x == 0 && it < 16;
              long long* pd = a.dbg + ((size_t)(it & 15) * 5 + s) * 8;
              if (probe) pd[5] = clock64();
              ch[p]++;
#if C2_FENCE_BY_ISSUER
              ptx::fence_proxy_async_smem();
#endif
              ptx::tc_fence_after_sync();
              issue(buf_u32 + (uint32_t)(2 + p) * (uint32_t)(C2_BUF * 16), w_u32, s, false);
              ptx::umma_commit(&ctl->mma_done[p]);
              if (probe) pd[6] = clock64();
              if (probe && s == nsteps) {  // how long until this hop's UMMAs are complete
                ptx::mbar_wait(&ctl->mma_done[p], (ch[p] - 1) & 1);
                pd[7] = clock64();
              }
              if (s == 1) ptx::umma_commit(&ctl->in_empty[st]);
            }
            ptx::umma_commit(&ctl->item_done[st]);
            if (c == n_chunks - 1) {
              ptx::umma_commit(&ctl->acc_full);
              ++g;
            }
            ++it;
          }
        }
      }
    }
    __syncwarp();
#endif
  } else {
    // ================================ input gather warps (+ accumulator drain, C2_IO_DRAIN) ================================
    C2_SETMAXNREG_DEC(C2_REG_IO);
    const int t = tid - (C2_NCW + 1) * 32;  // 0..95
    constexpr bool gathers = true;
    const int q = t & 1, pl0 = t >> 1;    // channel quad; position within a pair of lattice rows (0..47)
    const int inpos = pl0 % C2_LW, r0 = pl0 / C2_LW;
    const int col = col_of_slot(inpos);
    const int FV = a.F / 4;
    uint32_t it = 0;
#if C2_IO_OUT
    // Basis stores from the exchange buffer.  Unit u = t + 96 k (k < 6, u < 512): channel quad u & 1 (two consecutive
    // lanes write the 32 contiguous bytes of a pixel's chunk), centre position u >> 1 = 16 jj + cc.  The copies of item
    // i - 1 are served after the gather of item i (prefetch distance one item); `prev_*` is that item.
    const bool io_out = a.out[0] != nullptr || a.out[1] != nullptr || a.out[2] != nullptr || a.out[3] != nullptr;
    uint32_t chh[2] = {0, 0};
    int prev_rows[6] = {-1, -1, -1, -1, -1, -1};
    int64_t prev_b = 0;
    int prev_c = 0;
    bool prev_valid = false;
    int xoff[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const int u = t + C2_NLOAD * k, pi = (u >> 1) & 255;
      xoff[k] = (u & 1) * C2_PL + (C2_H + pi / C2_T + 1) * C2_LW + slot_of_col(C2_H + pi % C2_T);
    }
    auto serve = [&]() {  // all hops of the previous item: wait for the hop, copy its centre to out[s - 1], release X[p]
      for (int s = 1; s <= nsteps; ++s) {
        const int p = (s - 1) & 1;
        ptx::mbar_wait_backoff(&ctl->hop_full[p], chh[p] & 1, (uint32_t)a.sleep_ld);
        chh[p]++;
        float* outp = s == 1 ? a.out[0] : (s == 2 ? a.out[1] : (s == 3 ? a.out[2] : a.out[3]));
        if (outp != nullptr) {
          const float4* X = bufs + (size_t)(2 + p) * C2_BUF;
          float4* ob = reinterpret_cast<float4*>(outp + (prev_b * a.M * a.F + prev_c * C2_FC)) + q;
#pragma unroll
          for (int k = 0; k < 6; ++k) {
            const int row = prev_rows[k];
            if (row >= 0) __stcs(ob + (int64_t)row * FV, X[xoff[k]]);
          }
        }
        ptx::mbar_arrive(&ctl->out_done[p]);
      }
    };
#endif
#if C2_IO_DRAIN
    // groups (tile, b) whose last chunk has been passed, waiting for their drain: slot = group index & 3.  A group is
    // drained two items after its closing item: by then the next two items are gathered (the compute warps never wait
    // for this warp), and the wait for acc_full ends when the compute warps finish the closing item.
    uint32_t g_closed = 0, g_drained = 0;
    uint32_t close_it[4];
    int64_t close_b[4];
    int close_rows[4][3];
    auto drain = [&](uint32_t g) {
      const int sl = g & 3;
      const int rws[3] = {close_rows[sl][0], close_rows[sl][1], close_rows[sl][2]};
      ptx::mbar_wait_backoff(&ctl->acc_full, g & 1, (uint32_t)a.sleep_ld);
      ptx::tc_fence_after_sync();
      drain_quarter(a, tmem_base, warp, rws, close_b[sl]);
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&ctl->acc_empty);
    };
#endif
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      const int tile = unit / a.b_split;
      const int64_t b_begin = (int64_t)(unit % a.b_split) * b_per;
      const int64_t b_end = min(a.B, b_begin + b_per);
      if (b_begin >= b_end) continue;
      int rows[C2_LW / 2];
#pragma unroll
      for (int k = 0; k < C2_LW / 2; ++k) rows[k] = __ldg(a.pix + (size_t)tile * C2_P + (2 * k + r0) * C2_LW + col);
#if C2_IO_OUT
      int orow[6];  // rows of the basis tensor of this thread's 6 copy units (own pixels of this tile)
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int u = t + C2_NLOAD * k, pi = u >> 1;
        orow[k] = (io_out && u < 2 * C2_T * C2_T)
                      ? __ldg(a.pix + (size_t)tile * C2_P + (C2_H + pi / C2_T) * C2_LW + C2_H + pi % C2_T) : -1;
      }
#endif
#if C2_IO_DRAIN
      int erow[3];
#pragma unroll
      for (int mt = 0; mt < 3; ++mt) {  // accumulator row (mt, TMEM lane) -> lattice position -> row of y
        const int m = mt * 128 + (warp & 3) * 32 + lane;
        const int j = C2_H + m / C2_LW, p = m % C2_LW, c = col_of_slot(p);
        erow[mt] = (c >= C2_H && c < C2_H + C2_T) ? __ldg(a.pix + (size_t)tile * C2_P + j * C2_LW + c) : -1;
      }
#endif
      for (int64_t b = b_begin; b < b_end; ++b) {
        for (int c = 0; c < n_chunks; ++c) {
          const uint32_t st = it & 1;
          if (gathers) {
            if (it >= 2) ptx::mbar_wait_backoff(&ctl->in_empty[st], ((it >> 1) - 1) & 1, (uint32_t)a.sleep_ld);
            float4* dst = bufs + (size_t)st * C2_BUF + q * C2_PL + (r0 + 1) * C2_LW + inpos;
            const float4* src = reinterpret_cast<const float4*>(a.in0 + (b * a.M * a.F + c * C2_FC)) + q;
#pragma unroll
            for (int k = 0; k < C2_LW / 2; ++k) {
              const int row = rows[k];
              cp_async16(dst + 2 * k * C2_LW, row >= 0 ? (const void*)(src + (int64_t)row * FV) : (const void*)a.in0,
                         row >= 0 ? 16u : 0u);
            }
            cp_async_wait_all();
            ptx::fence_proxy_async_smem();
            ptx::mbar_arrive(&ctl->in_full[st]);
          }
#if C2_IO_OUT
          if (io_out) {
            if (prev_valid) serve();
#pragma unroll
            for (int k = 0; k < 6; ++k) prev_rows[k] = orow[k];
            prev_b = b; prev_c = c; prev_valid = true;
          }
#endif
#if C2_IO_DRAIN
          if (c == n_chunks - 1) {
            const int sl = g_closed & 3;
            close_it[sl] = it; close_b[sl] = b;
            close_rows[sl][0] = erow[0]; close_rows[sl][1] = erow[1]; close_rows[sl][2] = erow[2];
            ++g_closed;
          }
          while (g_drained < g_closed && close_it[g_drained & 3] + 2 <= it) drain(g_drained++);
#endif
          ++it;
        }
      }
    }
#if C2_IO_DRAIN
    while (g_drained < g_closed) drain(g_drained++);
#endif
#if C2_IO_OUT
    if (io_out && prev_valid) serve();
#endif
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == C2_NCW) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, C2_TMEM_COLS);
  }
}

// weight images: chunk c, hop k: element (n, kk) at (kk / 4) * (N * 4) + n * 4 + kk % 4,
// value = tf32_rn( Wsrc[(c*16 + kk) * s_f + k * s_k + n * s_n] )
__global__ void conv2_prep_b_kernel(const float* __restrict__ W, int64_t s_f, int64_t s_k, int64_t s_n, int n_chunks,
                                    int K, int N, float* __restrict__ img) {
  const int64_t per = (int64_t)N * C2_FC;
  const int64_t total = (int64_t)n_chunks * K * per;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int kk = (int)(e % C2_FC);
    const int n = (int)((e / C2_FC) % N);
    const int k = (int)((e / per) % K);
    const int c = (int)(e / (per * K));
    // odd hops leave T_k in shared memory with each channel quad rotated by one (c1, c2, c3, c0)
    const int ksrc = hop_is_rotated(k) ? ((kk & ~3) | ((kk + 1) & 3)) : kk;
    const float v = W[(int64_t)(c * C2_FC + ksrc) * s_f + (int64_t)k * s_k + (int64_t)n * s_n];
    uint32_t u = __float_as_uint(v);
    uint32_t rr = (u + 0x00000FFFu + ((u >> 13) & 1u)) & 0xFFFFE000u;
    float hi = __uint_as_float(rr);
    if (!isfinite(hi)) hi = __uint_as_float(u & 0xFFFFE000u);
    img[((int64_t)(c * K + k)) * per + (int64_t)(kk / 4) * (N * 4) + n * 4 + (kk % 4)] = hi;
  }
}

size_t conv2_smem_bytes(int N, int nsteps) {
  return (size_t)4 * C2_BUF * 16 + (size_t)2 * (nsteps + 1) * N * C2_FC * 4 + (size_t)C2_P * 4 + sizeof(Conv2Ctl) + 16;
}

}  // namespace

#ifndef DS_EMULATE  // tests/emul runs the kernels above on the host; the launchers below need nvcc

// Is the register-resident fused kernel available for this call?
bool lattice_conv2_usable(const LatticeDev& L, int nsteps, int F, int N, int mode) {
  if (mode != DS_MODE_TF32 || L.n_tiles <= 0) return false;
  if (L.T != C2_T || L.H != C2_H || L.LW != C2_LW) return false;
  if (C2_CDIAG && !L.diag_const) return false;  // this build keeps the diagonal of L~ as one scalar
  if (nsteps < 1 || nsteps > C2_H) return false;
  if (F % C2_FC != 0 || N % 16 != 0 || N < 16 || 3 * N > C2_TMEM_COLS) return false;
  static const bool disabled = [] { const char* e = getenv("DEEPSPHERE_FUSED_CONV2"); return e && atoi(e) == 0; }();
  if (disabled) return false;
  int dev = 0, max_smem = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  if (cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return false;
  return (int)conv2_smem_bytes(N, nsteps) <= max_smem;
}

// y = act( sum_k T_k(L~)(in0) * B_k + bias ) on the regular tiles; out[s-1] (optional) receives T_s(in0) (own pixels).
// Weights are addressed generically: B_k(f, n) = W[f*s_f + k*s_k + n*s_n].
int launch_lattice_conv2(const LatticeDev& L, int nsteps, int64_t B, int64_t M, int F, int N, int recursion,
                         const float* in0, float* const* out, const float* W, int64_t s_f, int64_t s_k, int64_t s_n,
                         const float* bias, int act, float* y, cudaStream_t st) {
  const bool cheb = recursion == DS_RECURSION_CHEBYSHEV;
  const int K = nsteps + 1, n_chunks = F / C2_FC;
  Conv2Args a;
  a.n_tiles = L.n_tiles; a.pix = L.pix; a.w = L.w;
  a.B = B; a.M = M; a.F = F; a.N = N; a.nsteps = nsteps;
  int split = 1;  // persistent CTAs, 2 per SM: enough units for a balanced tail
  while ((int64_t)L.n_tiles * split < (int64_t)32 * num_sms() && split < B) split *= 2;
  // sweeps without a rebuild: DEEPSPHERE_CONV2_BSPLIT (work units per tile along the batch), DEEPSPHERE_CONV2_GRID (CTAs)
  static const int env_split = [] { const char* e = getenv("DEEPSPHERE_CONV2_BSPLIT"); return e ? atoi(e) : 0; }();
  static const int env_grid = [] { const char* e = getenv("DEEPSPHERE_CONV2_GRID"); return e ? atoi(e) : 0; }();
  if (env_split > 0) split = env_split;
  a.b_split = (int)std::min<int64_t>(split, B);
  a.wscale = cheb ? 2.f : 1.f;
  a.wdiag = a.wscale * L.diag;
  static const int sleep_mma = [] { const char* e = getenv("DEEPSPHERE_CONV2_SLEEP_MMA"); return e ? atoi(e) : 0; }();
  static const int sleep_ld = [] { const char* e = getenv("DEEPSPHERE_CONV2_SLEEP_LD"); return e ? atoi(e) : 128; }();
  a.sleep_mma = sleep_mma; a.sleep_ld = sleep_ld;
  static const bool dbg_on = [] { const char* e = getenv("DEEPSPHERE_CONV2_DEBUG"); return e && atoi(e) == 1; }();
  a.dbg = nullptr;
  if (dbg_on) {
    DS_CUDA(cudaMalloc((void**)&a.dbg, 16 * 5 * 8 * sizeof(long long)));
    DS_CUDA(cudaMemsetAsync(a.dbg, 0, 16 * 5 * 8 * sizeof(long long), st));
  }
  for (int s = 0; s < C2_H; ++s) a.out[s] = (out != nullptr && s < nsteps) ? out[s] : nullptr;
  a.in0 = in0; a.bias = bias; a.act = act; a.y = y;
  float* img = nullptr;
  const size_t img_elems = (size_t)n_chunks * K * N * C2_FC;
  DS_CUDA(cudaMallocAsync((void**)&img, img_elems * 4, st));
  conv2_prep_b_kernel<<<(unsigned)std::min<size_t>((img_elems + 255) / 256, 1024), 256, 0, st>>>(
      W, s_f, s_k, s_n, n_chunks, K, N, img);
  g_launches.fetch_add(1);
  a.b_img = img;
  const size_t smem = conv2_smem_bytes(N, nsteps);
  static PerDeviceOnce attr_once;
  const int attr_rc = attr_once.run([&]() -> int {
    int dev = 0, max_smem = 0;
    DS_CUDA(cudaGetDevice(&dev));
    DS_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    DS_CUDA(cudaFuncSetAttribute(lattice_conv2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    DS_CUDA(cudaFuncSetAttribute(lattice_conv2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    DS_CUDA(cudaFuncSetAttribute(lattice_conv2_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    DS_CUDA(cudaFuncSetAttribute(lattice_conv2_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    return 0;
  });
  if (attr_rc != 0) {
    cudaFreeAsync(img, st);
    if (a.dbg != nullptr) cudaFree(a.dbg);
    return attr_rc;
  }
  const int n_units = a.n_tiles * a.b_split;
  const int grid = std::min(n_units, env_grid > 0 ? env_grid : 2 * num_sms());
  if (cheb) lattice_conv2_kernel<true><<<grid, C2_THREADS, smem, st>>>(a);
  else lattice_conv2_kernel<false><<<grid, C2_THREADS, smem, st>>>(a);
  cudaError_t e = cudaGetLastError();
  if (a.dbg != nullptr) {
    long long h[16 * 5 * 8];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, a.dbg, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(a.dbg);
    const long long t0 = h[0];
    for (int it = 0; it < 16; ++it) {
      fprintf(stderr, "item %2d: wait_in %lld..%lld\n", it, h[(it * 5) * 8] - t0, h[(it * 5) * 8 + 1] - t0);
      for (int s2 = 1; s2 <= nsteps; ++s2) {
        const long long* q = h + (it * 5 + s2) * 8;
        fprintf(stderr, "   hop %d: start %lld fma+%lld mmawait+%lld stsfence+%lld bar+%lld | mma: seen %lld issued+%lld done+%lld\n", s2,
                q[0] - t0, q[1] - q[0], q[2] - q[1], q[3] - q[2], q[4] - q[3], q[5] - t0, q[6] - q[5], q[7] ? q[7] - q[6] : 0);
      }
    }
  }
  cudaFreeAsync(img, st);
  g_launches.fetch_add(1);
  if (e != cudaSuccess) return fail("lattice_conv2_kernel launch failed: %s", cudaGetErrorString(e));
  return 0;
}

#endif  // DS_EMULATE

}  // namespace ds
