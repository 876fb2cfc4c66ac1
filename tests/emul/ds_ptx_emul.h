// Host emulation of csrc/ds_ptx.cuh — TEST INFRASTRUCTURE (selected by -DDS_EMULATE).  Functional models of the sm_100a
// features the kernels use, so that an UNCHANGED kernel source runs on CPU threads:
//   * mbarrier: arrival count + transaction bytes + phase, each barrier behind its own mutex / condition variable
//   * cp.async.bulk (1-D): memcpy + complete_tx
//   * tcgen05: tensor memory = 128 lanes x 512 columns of 32 bits per CTA; tcgen05.mma.kind::tf32 decodes the shared
//     memory descriptors (no-swizzle K-major core matrices: 8 rows x 16 bytes; LBO = step between core matrices along
//     K, SBO = step along M / N), truncates the operands to TF32 like the hardware and accumulates in fp32;
//     tcgen05.commit arrives at once (emulated UMMAs complete synchronously); tcgen05.ld 32x32b.x16
//   * proxy / tcgen05 fences: no-ops (every emulated access is sequentially consistent)
// What it cannot show: anything about timing, and memory-model violations that only a weaker ordering exposes.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ds {
namespace ptx {

inline uint32_t smem_u32(const void* p) {
  return (uint32_t)(reinterpret_cast<const uint8_t*>(p) - emul::state()->dyn_smem);
}
inline uint8_t* smem_ptr(uint32_t addr) { return emul::state()->dyn_smem + addr; }

// ---- mbarrier ---------------------------------------------------------------------------
inline emul::MbarState& mbar_slot(const void* bar) {
  emul::BlockState* s = emul::state();
  const size_t off = (size_t)(reinterpret_cast<const uint8_t*>(bar) - s->dyn_smem);
  if (off % 8 != 0 || off >= emul::kSmemBytes) { std::fprintf(stderr, "emul: mbarrier outside dynamic shared memory\n"); std::abort(); }
  return s->mbar[off / 8];
}
inline emul::MbarState& mbar_live(const void* bar) {
  emul::MbarState& m = mbar_slot(bar);
  if (!m.live) { std::fprintf(stderr, "emul: use of an uninitialised mbarrier\n"); std::abort(); }
  return m;
}
// caller holds m.mu
inline void mbar_check_locked(emul::MbarState& m) {
  if (m.pending < 0) { std::fprintf(stderr, "emul: mbarrier over-arrived\n"); std::abort(); }
  if (m.pending == 0 && m.tx == 0) {
    m.phase++;
    m.pending = (int32_t)m.expected;
    m.cv.notify_all();
  }
}
inline void mbar_init(uint64_t* bar, uint32_t count) {
  emul::MbarState& m = mbar_slot(bar);
  std::lock_guard<std::mutex> lk(m.mu);
  m.live = true; m.expected = count; m.pending = (int32_t)count; m.tx = 0; m.phase = 0;
}
inline void fence_mbar_init() {}
inline void mbar_arrive(uint64_t* bar) {
  emul::MbarState& m = mbar_live(bar);
  std::lock_guard<std::mutex> lk(m.mu);
  m.pending -= 1;
  mbar_check_locked(m);
}
inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  emul::MbarState& m = mbar_live(bar);
  std::lock_guard<std::mutex> lk(m.mu);
  m.tx += bytes;
  m.pending -= 1;
  mbar_check_locked(m);
}
inline void mbar_complete_tx(uint64_t* bar, uint32_t bytes) {
  emul::MbarState& m = mbar_live(bar);
  std::lock_guard<std::mutex> lk(m.mu);
  m.tx -= bytes;
  mbar_check_locked(m);
}
// the phase with parity `parity` has completed <=> the barrier's current phase has the other parity
inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  emul::MbarState& m = mbar_live(bar);
  std::unique_lock<std::mutex> lk(m.mu);
  return m.cv.wait_for(lk, std::chrono::milliseconds(2), [&] { return (m.phase & 1u) != parity; });
}
inline bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  emul::MbarState& m = mbar_live(bar);
  std::lock_guard<std::mutex> lk(m.mu);
  return (m.phase & 1u) != parity;
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  const auto t0 = std::chrono::steady_clock::now();
  while (!mbar_try_wait(bar, parity)) {
    if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) {
      std::fprintf(stderr, "emul: mbarrier wait timed out (block %u thread %u, barrier at shared offset %td, parity %u) - protocol deadlock\n",
                   blockIdx.x, threadIdx.x, reinterpret_cast<uint8_t*>(bar) - emul::dynamic_smem(), parity);
      std::abort();
    }
  }
}
inline void mbar_wait_backoff(uint64_t* bar, uint32_t parity, uint32_t) { mbar_wait(bar, parity); }
inline void st_async_v4(void* dst, const float4& v, uint64_t* bar) {
  *reinterpret_cast<float4*>(dst) = v;
  mbar_complete_tx(bar, 16);
}

// ---- proxies / fences ---------------------------------------------------------------------
inline void fence_proxy_async_smem() {}
inline void tc_fence_before_sync() {}
inline void tc_fence_after_sync() {}

inline void named_bar_sync(int id, int nthreads) {
  emul::BlockState* s = emul::state();
  std::barrier<>* b;
  {
    std::lock_guard<std::mutex> lk(s->named_mu);
    auto& slot = s->named[id];
    if (!slot) slot = std::make_unique<std::barrier<>>(nthreads);
    b = slot.get();
  }
  b->arrive_and_wait();
}
inline void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  std::memcpy(dst, src, bytes);
  mbar_complete_tx(bar, bytes);
}

// ---- tensor memory --------------------------------------------------------------------------
inline void tmem_alloc(uint32_t* smem_dst, uint32_t) {  // warp-collective: one write
  if ((threadIdx.x & 31) == 0) *smem_dst = 0;
}
inline void tmem_dealloc(uint32_t, uint32_t) {}

constexpr uint64_t LAYOUT_SWIZZLE_NONE = 0, LAYOUT_SWIZZLE_128B_BASE32B = 1, LAYOUT_SWIZZLE_128B = 2, LAYOUT_SWIZZLE_64B = 4, LAYOUT_SWIZZLE_32B = 6;
inline uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint64_t layout_type) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (layout_type << 61);
}
constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a_mn_major & 1) << 15) | ((uint32_t)(b_mn_major & 1) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
inline float tf32_trunc(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
// element (row r, k) of a K-major, no-swizzle operand: core matrix = 8 rows x 16 bytes
inline float umma_operand(uint64_t desc, int r, int k) {
  const uint32_t start = (uint32_t)(desc & 0x3FFF) << 4, lbo = (uint32_t)((desc >> 16) & 0x3FFF) << 4,
                 sbo = (uint32_t)((desc >> 32) & 0x3FFF) << 4;
  if ((desc >> 61) != LAYOUT_SWIZZLE_NONE) { std::fprintf(stderr, "emul: only no-swizzle UMMA operands are modelled\n"); std::abort(); }
  const uint32_t addr = start + (uint32_t)(r / 8) * sbo + (uint32_t)(k / 4) * lbo + (uint32_t)(r % 8) * 16 + (uint32_t)(k % 4) * 4;
  float v;
  std::memcpy(&v, smem_ptr(addr), 4);
  return v;
}
// D[tmem] (+)= A[smem] * B[smem]^T, one K step of 8 (tf32), M = 128 rows on lanes 0..127
inline void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  const int M = (int)((idesc >> 24) & 0x1F) << 4, N = (int)((idesc >> 17) & 0x3F) << 3;
  if (M != 128 || ((idesc >> 15) & 3) != 0) { std::fprintf(stderr, "emul: UMMA shape / major-ness not modelled\n"); std::abort(); }
  uint32_t* tm = emul::state()->tmem.data();
  const uint32_t lane0 = d_tmem >> 16, col0 = d_tmem & 0xFFFF;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0.0;
      for (int k = 0; k < 8; ++k) acc += (double)tf32_trunc(umma_operand(a_desc, m, k)) * (double)tf32_trunc(umma_operand(b_desc, n, k));
      uint32_t& cell = tm[(size_t)(lane0 + m) * 512 + col0 + n];
      const float prev = accumulate ? __uint_as_float(cell) : 0.f;
      cell = __float_as_uint((float)((double)prev + acc));
    }
}
inline void umma_commit(uint64_t* bar) { mbar_arrive(bar); }
inline void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  const uint32_t* tm = emul::state()->tmem.data();
  const uint32_t lane = (taddr >> 16) + (threadIdx.x & 31), col0 = taddr & 0xFFFF;
  for (int i = 0; i < 16; ++i) r[i] = tm[(size_t)lane * 512 + col0 + i];
}
inline void tmem_ld_wait() {}

}  // namespace ptx
}  // namespace ds
