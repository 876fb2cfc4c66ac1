// CPU stand-in for <cuda_runtime.h> — TEST INFRASTRUCTURE.  Lets g++ compile a simple CUDA kernel source unchanged and
// run its thread blocks on the host (tests/emul/README in cuda_emul.h): every CUDA thread of a block is a std::thread,
// __syncthreads() is a block-wide barrier, warp shuffles go through a per-warp exchange buffer.  Only what the
// emulated sources (csrc/ds_common.cuh, csrc/ds_skinny.cu) use is provided.
#pragma once
#include <barrier>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)
#define __noinline__

struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
struct int4 { int x, y, z, w; };
struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline float2 make_float2(float x, float y) { return float2{x, y}; }
// packed fp32 FMA: two IEEE fused multiply-adds
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)}; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline long long clock64() { return (long long)std::chrono::steady_clock::now().time_since_epoch().count(); }
inline void __trap() { std::fprintf(stderr, "emul: __trap()\n"); std::abort(); }
inline void __nanosleep(unsigned) { std::this_thread::yield(); }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long min(long a, long b) { return a < b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline long max(long a, long b) { return a > b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
using std::isfinite;

typedef int cudaError_t;
typedef void* cudaStream_t;
constexpr cudaError_t cudaSuccess = 0;
constexpr int cudaDevAttrMultiProcessorCount = 16;
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 148; return cudaSuccess; }

namespace emul {
// one synchronisation object PER mbarrier (not a global lock): ThreadSanitizer then only sees the happens-before edges
// the barrier protocol really creates, so a missing wait between two roles shows up as a data race
struct MbarState {
  std::mutex mu;
  std::condition_variable cv;
  bool live = false;
  uint32_t expected = 0;
  int32_t pending = 0;
  int64_t tx = 0;
  uint32_t phase = 0;
};
constexpr size_t kSmemBytes = 232448;  // 227 KB
struct BlockState {
  std::unique_ptr<std::barrier<>> block_bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<float> shfl;  // [warps][32]
  // dynamic shared memory, tensor memory (128 lanes x 512 columns), mbarriers and named barriers of the block
  std::vector<uint8_t> dyn_smem_storage;
  uint8_t* dyn_smem = nullptr;
  std::vector<uint32_t> tmem;
  std::unique_ptr<MbarState[]> mbar;  // indexed by the barrier's 8-byte slot in dynamic shared memory
  std::mutex named_mu;
  std::map<int, std::unique_ptr<std::barrier<>>> named;
};
inline BlockState*& state() { static BlockState* s = nullptr; return s; }
}  // namespace emul

inline thread_local uint3 threadIdx{0, 0, 0};
inline thread_local uint3 blockIdx{0, 0, 0};
inline thread_local dim3 blockDim;
inline thread_local dim3 gridDim;

namespace emul {
inline uint8_t* dynamic_smem() { return state()->dyn_smem; }
}  // namespace emul
inline void __syncthreads() { emul::state()->block_bar->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emul::state()->warp_bar[threadIdx.x >> 5]->arrive_and_wait(); }
template <class T>
inline void __stcs(T* p, const T& v) { *p = v; }
inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
  emul::BlockState* s = emul::state();
  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  s->shfl[warp * 32 + lane] = v;
  s->warp_bar[warp]->arrive_and_wait();
  const float r = s->shfl[warp * 32 + (lane ^ (unsigned)lane_mask)];
  s->warp_bar[warp]->arrive_and_wait();
  return r;
}
template <class T>
inline T __ldg(const T* p) { return *p; }

namespace emul {
// run `kernel()` for every thread of a grid x block launch (1-D), blocks one after the other
inline void launch(unsigned grid, unsigned block, const std::function<void()>& kernel, size_t dyn_smem_bytes = 0) {
  if (block % 32 != 0) { std::fprintf(stderr, "emul: block size must be a multiple of 32\n"); std::abort(); }
  for (unsigned b = 0; b < grid; ++b) {
    BlockState st;
    st.dyn_smem_storage.assign(dyn_smem_bytes + 1024, 0xCD);  // poisoned: reads of unwritten smem show up
    st.dyn_smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(st.dyn_smem_storage.data()) + 1023) & ~uintptr_t(1023));
    st.tmem.assign(128 * 512, 0xCDCDCDCDu);
    st.mbar = std::make_unique<MbarState[]>(kSmemBytes / 8);
    st.block_bar = std::make_unique<std::barrier<>>(block);
    for (unsigned w = 0; w < block / 32; ++w) st.warp_bar.push_back(std::make_unique<std::barrier<>>(32));
    st.shfl.assign(block, 0.f);
    state() = &st;
    std::vector<std::thread> threads;
    for (unsigned t = 0; t < block; ++t)
      threads.emplace_back([&, t] {
        threadIdx = uint3{t, 0, 0};
        blockIdx = uint3{b, 0, 0};
        blockDim = dim3(block);
        gridDim = dim3(grid);
        kernel();
      });
    for (auto& th : threads) th.join();
    state() = nullptr;
  }
}
}  // namespace emul
