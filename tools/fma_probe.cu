// Micro-probe: fp32 FFMA vs packed FFMA2 issue rate per SM sub-partition on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fma_probe tools/fma_probe.cu  (binary is git-ignored)
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void probe(float* out, int iters, long long* cycles) {
  float2 acc[16];
  float w[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = 1.0f + 1e-6f * (threadIdx.x + i);
  float2 x = make_float2(0.999f, 1.001f);
  float2 xs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) xs[i] = make_float2(0.999f + 1e-4f * i + 1e-6f * threadIdx.x, 1.001f - 1e-4f * i);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (MODE == 0) {  // scalar FFMA, 2 per accumulator pair
          acc[i].x = fmaf(w[j], x.x, acc[i].x);
          acc[i].y = fmaf(w[j], x.y, acc[i].y);
        } else if (MODE == 1) {  // FFMA2, broadcast pair built once per j (hoisted)
          float2 ww = make_float2(w[j], w[j]);
          acc[i] = __ffma2_rn(ww, x, acc[i]);
        } else if (MODE == 3) {  // FFMA2, scalar weight reused across i, data pair distinct per instruction
          acc[i] = __ffma2_rn(make_float2(w[j], w[j]), xs[(i + j) & 7], acc[i]);
        } else if (MODE == 4) {  // FFMA2, weight AND data distinct per instruction
          acc[i] = __ffma2_rn(make_float2(w[(i + j) & 7], w[(i + j) & 7]), xs[(i + 3 * j) & 7], acc[i]);
        } else if (MODE == 5) {  // scalar FFMA, weight and data distinct per instruction
          acc[i].x = fmaf(w[(i + j) & 7], xs[(i + 3 * j) & 7].x, acc[i].x);
          acc[i].y = fmaf(w[(i + j) & 7], xs[(i + 3 * j) & 7].y, acc[i].y);
        } else {  // FFMA2 with a fresh pair per 2 FFMA2 (forces a MOV per pair of FFMA2)
          float wv = w[j] + (float)(i >> 1) * 1e-9f * (float)it;
          float2 ww = make_float2(wv, wv);
          acc[i] = __ffma2_rn(ww, x, acc[i]);
        }
      }
    }
    x.x += 1e-7f;
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps_per_sm) {
  int threads = warps_per_sm * 32, blocks = 148, iters = 2000;
  float* out; long long* cyc;
  cudaMalloc(&out, sizeof(float) * threads * blocks);
  cudaMalloc(&cyc, sizeof(long long) * blocks);
  probe<MODE><<<blocks, threads>>>(out, iters, cyc);
  probe<MODE><<<blocks, threads>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double fmas = (double)iters * 8 * 16 * 2 * threads;  // per SM
  printf("%-28s warps/SM %2d: %8lld cycles, %.1f FMA/clk/SM  (%s)\n", name, warps_per_sm, h[0], fmas / (double)h[0],
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<0>("FFMA scalar", w);
    run<1>("FFMA2 (pair hoisted)", w);
    run<2>("FFMA2 + op per pair", w);
    run<3>("FFMA2 w reused, x distinct", w);
    run<4>("FFMA2 w, x distinct", w);
    run<5>("FFMA w, x distinct", w);
  }
  return 0;
}
