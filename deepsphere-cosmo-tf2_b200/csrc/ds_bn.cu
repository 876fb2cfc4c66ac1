// BatchNormalization(axis=-1, momentum, epsilon, center=False, scale=False) + bias + activation around the graph
// convolution (reference gnn_layers.py:53, 152-159: `x = self.bn(x, training)`, `x = tf.add(x, self.bias)`,
// `x = self.activation(x)`), forward and backward, as HBM-bound streaming kernels on the native [B, M, F] layout.
//
//   forward   ds_bn_stats            per-channel sum z, sum z^2 over rows [r0, r1) of every sample  -> double[2F]
//             (caller: all-reduce of the 2F+1 numbers over the ranks for the statistics of the GLOBAL batch)
//             ds_bn_bias_act_forward finalise (mean, 1/sqrt(var + eps), moving statistics) + y = act((z - mean) rstd + bias)
//   backward  ds_bn_backward_stats   g = dy * act'(y); per-channel sum g, sum g * zhat over the same rows   -> double[2F]
//             (caller: all-reduce)
//             ds_bn_backward_apply   dz = rstd (g - mean(g) - zhat mean(g zhat)) inside [r0, r1), 0 outside; dbias = sum g
//
// The row range serves the sphere-partitioned layers: the statistics (and the loss) only see a rank's OWN rows, the
// halo rows of the extended set are normalised with the same statistics and carry no gradient.
// Reductions: fp32 per-thread partial sums over <= a few thousand elements, combined in double, block partials written
// to a workspace and summed in a fixed order by a second kernel (bit-reproducible; E[z^2] - E[z]^2 is formed in double).
#include <algorithm>

#include "ds_common.cuh"

namespace ds {
namespace {

constexpr int BN_THREADS = 256;

// number of blocks: threads * blocks must be a multiple of F so that a thread always sees the same channel
inline int bn_blocks(int64_t n_elems, int64_t F) {
  int64_t g = 1;
  {  // gcd(F, BN_THREADS)
    int64_t a = F, b = BN_THREADS;
    while (b) { const int64_t t = a % b; a = b; b = t; }
    g = a;
  }
  const int64_t unit = F / g;  // blocks must be a multiple of this
  int64_t want = std::min<int64_t>((n_elems + BN_THREADS * 8 - 1) / (BN_THREADS * 8), (int64_t)num_sms() * 8);
  want = std::max<int64_t>(1, want);
  int64_t blocks = (want + unit - 1) / unit * unit;
  return (int)blocks;
}

// element e of the restricted index space (b, r - r0, f) -> offset in the [B, M, F] tensor
__device__ __forceinline__ int64_t bn_offset(int64_t e, int64_t span, int64_t M, int64_t F, int64_t r0) {
  const int64_t b = e / span, rem = e - b * span;
  return (b * M + r0) * F + rem;
}

// MODE 0: (z, z^2);  MODE 1: (g, g * zhat) with g = dy * act'(y), zhat = (z - mean) * rstd
template <int MODE>
__global__ void __launch_bounds__(BN_THREADS) bn_partial_kernel(int64_t B, int64_t M, int64_t F, int64_t r0, int64_t r1,
                                                                const float* __restrict__ z, const float* __restrict__ y,
                                                                const float* __restrict__ dy,
                                                                const float* __restrict__ mean_rstd, int act,
                                                                double* __restrict__ partial) {
  __shared__ double s_a[BN_THREADS], s_b[BN_THREADS];
  const int64_t span = (r1 - r0) * F, total = B * span;
  const int64_t T = (int64_t)gridDim.x * BN_THREADS, t0 = (int64_t)blockIdx.x * BN_THREADS + threadIdx.x;
  const int c = (int)(t0 % F);
  float mean = 0.f, rstd = 1.f;
  if (MODE == 1) { mean = mean_rstd[c]; rstd = mean_rstd[F + c]; }
  double da = 0.0, db = 0.0;
  float a = 0.f, b = 0.f;
  int cnt = 0;
  for (int64_t e = t0; e < total; e += T) {
    const int64_t off = bn_offset(e, span, M, F, r0);
    if (MODE == 0) {
      const float v = __ldg(z + off);
      a += v;
      b = fmaf(v, v, b);
    } else {
      float g = __ldg(dy + off);
      if (act != DS_ACT_LINEAR) g *= act_grad_from_y(__ldg(y + off), act);
      a += g;
      b = fmaf(g, (__ldg(z + off) - mean) * rstd, b);
    }
    if (++cnt == 1024) {  // bound the fp32 accumulation length
      da += a; db += b; a = b = 0.f; cnt = 0;
    }
  }
  s_a[threadIdx.x] = da + a;
  s_b[threadIdx.x] = db + b;
  __syncthreads();
  // channel of thread t in this block: (block_base + t) % F; reducer thread j handles channels j, j + 256, ...
  const int64_t base = (int64_t)blockIdx.x * BN_THREADS;
  for (int64_t ch = threadIdx.x; ch < F; ch += BN_THREADS) {
    int64_t first = (ch - base % F + F) % F;  // first thread of the block with this channel
    double ra = 0.0, rb = 0.0;
    for (int64_t t = first; t < BN_THREADS; t += F) { ra += s_a[t]; rb += s_b[t]; }
    partial[((int64_t)blockIdx.x * 2) * F + ch] = ra;
    partial[((int64_t)blockIdx.x * 2 + 1) * F + ch] = rb;
  }
}

// sums[i] = sum over the blocks' partials, one WARP per output i (fixed lane assignment + shuffle tree: deterministic)
__global__ void __launch_bounds__(256) bn_reduce_kernel(int blocks, int64_t F, const double* __restrict__ partial,
                                                        double* __restrict__ sums) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);  // 0 .. 2F
  if (i >= 2 * F) return;
  const int64_t which = i / F, ch = i - which * F;
  double acc = 0.0;
  for (int k = lane; k < blocks; k += 32) acc += partial[((int64_t)k * 2 + which) * F + ch];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) sums[i] = acc;
}

// mean / rstd from the (all-reduced) sums, moving statistics, and the affine form of the normalisation
__global__ void bn_finalize_kernel(int64_t F, const double* __restrict__ sums, double count,
                                   const double* __restrict__ count_dev, float eps, float momentum,
                                   int training, float* __restrict__ moving_mean, float* __restrict__ moving_var,
                                   const float* __restrict__ bias, float* __restrict__ mean_rstd,
                                   float* __restrict__ scale_shift) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= F) return;
  float mean, var;
  if (training) {
    if (count_dev != nullptr) count = *count_dev;
    const double m = sums[c] / count;
    const double v = fmax(sums[F + c] / count - m * m, 0.0);  // biased variance, as Keras
    mean = (float)m;
    var = (float)v;
    moving_mean[c] = moving_mean[c] * momentum + mean * (1.f - momentum);
    moving_var[c] = moving_var[c] * momentum + var * (1.f - momentum);
  } else {
    mean = moving_mean[c];
    var = moving_var[c];
  }
  const float rstd = rsqrtf(var + eps);
  mean_rstd[c] = mean;
  mean_rstd[F + c] = rstd;
  scale_shift[c] = rstd;
  scale_shift[F + c] = (bias != nullptr ? bias[c] : 0.f) - mean * rstd;
}

// y = act(z * scale[c] + shift[c])
template <int V>
__global__ void __launch_bounds__(256) bn_apply_kernel(int64_t n_vec, int64_t FV, const float* __restrict__ z,
                                                       const float* __restrict__ scale_shift, int act,
                                                       float* __restrict__ y) {
  const int64_t F = FV * V;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_vec; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c0 = (e % FV) * V;
    if (V == 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(z) + e);
      const float4 sc = __ldg(reinterpret_cast<const float4*>(scale_shift + c0));
      const float4 sh = __ldg(reinterpret_cast<const float4*>(scale_shift + F + c0));
      float4 o = make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w));
      if (act != DS_ACT_LINEAR) o = make_float4(act_apply(o.x, act), act_apply(o.y, act), act_apply(o.z, act), act_apply(o.w, act));
      reinterpret_cast<float4*>(y)[e] = o;
    } else {
      const float o = fmaf(__ldg(z + e), __ldg(scale_shift + c0), __ldg(scale_shift + F + c0));
      y[e] = act != DS_ACT_LINEAR ? act_apply(o, act) : o;
    }
  }
}

// dz = rstd * (g - sum_g / n - zhat * sum_gz / n) for rows in [r0, r1), 0 elsewhere
__global__ void __launch_bounds__(256) bn_backward_apply_kernel(int64_t B, int64_t M, int64_t F, int64_t r0, int64_t r1,
                                                                const float* __restrict__ z, const float* __restrict__ y,
                                                                const float* __restrict__ dy,
                                                                const float* __restrict__ mean_rstd,
                                                                const double* __restrict__ sums, double count,
                                                                const double* __restrict__ count_dev, int act,
                                                                int training, float* __restrict__ dz) {
  const int64_t total = B * M * F;
  if (count_dev != nullptr) count = *count_dev;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = e % F, r = (e / F) % M;
    float out = 0.f;
    if (r >= r0 && r < r1) {
      float g = __ldg(dy + e);
      if (act != DS_ACT_LINEAR) g *= act_grad_from_y(__ldg(y + e), act);
      const float mean = mean_rstd[c], rstd = mean_rstd[F + c];
      if (training) {
        const float zh = (__ldg(z + e) - mean) * rstd;
        out = rstd * (g - (float)(sums[c] / count) - zh * (float)(sums[F + c] / count));
      } else {
        out = rstd * g;  // moving statistics are constants
      }
    }
    dz[e] = out;
  }
}

__global__ void bn_dbias_kernel(int64_t F, const double* __restrict__ sums, float* __restrict__ dbias) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < F) dbias[c] = (float)sums[c];
}

inline unsigned stream_grid(int64_t n) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 16));
}

template <int MODE>
int bn_sums(int64_t B, int64_t M, int64_t F, int64_t r0, int64_t r1, const float* z, const float* y, const float* dy,
            const float* mean_rstd, int act, double* sums, double* workspace, cudaStream_t st) {
  const int blocks = bn_blocks(B * (r1 - r0) * F, F);
  bn_partial_kernel<MODE><<<blocks, BN_THREADS, 0, st>>>(B, M, F, r0, r1, z, y, dy, mean_rstd, act, workspace);
  DS_LAUNCHED();
  bn_reduce_kernel<<<(unsigned)((2 * F + 7) / 8), 256, 0, st>>>(blocks, F, workspace, sums);
  DS_LAUNCHED();
  return 0;
}

}  // namespace
}  // namespace ds

extern "C" int64_t ds_bn_workspace_doubles(int64_t B, int64_t M, int64_t F) {
  if (B <= 0 || M <= 0 || F <= 0) return 0;
  return (int64_t)ds::bn_blocks(B * M * F, F) * 2 * F;
}

extern "C" int ds_bn_stats(int64_t B, int64_t M, int64_t F, int64_t r0, int64_t r1, const float* z, double* sums,
                           double* workspace, void* stream) {
  using namespace ds;
  DS_CHECK(B > 0 && M > 0 && F > 0 && r0 >= 0 && r0 < r1 && r1 <= M, "ds_bn_stats: bad sizes");
  DS_CHECK(z && sums && workspace, "ds_bn_stats: NULL pointer");
  return bn_sums<0>(B, M, F, r0, r1, z, nullptr, nullptr, nullptr, DS_ACT_LINEAR, sums, workspace, (cudaStream_t)stream);
}

extern "C" int ds_bn_bias_act_forward(int64_t B, int64_t M, int64_t F, const float* z, const double* sums, double count,
                                      const double* count_dev, float eps, float momentum, int32_t training, float* moving_mean, float* moving_var,
                                      const float* bias, int32_t act, float* mean_rstd, float* scale_shift, float* y,
                                      void* stream) {
  using namespace ds;
  DS_CHECK(B > 0 && M > 0 && F > 0, "ds_bn_bias_act_forward: bad sizes");
  DS_CHECK(z && y && moving_mean && moving_var && mean_rstd && scale_shift, "ds_bn_bias_act_forward: NULL pointer");
  DS_CHECK(!training || (sums != nullptr && (count > 0 || count_dev != nullptr)),
           "ds_bn_bias_act_forward: training needs the batch sums and their row count");
  DS_CHECK(act >= DS_ACT_LINEAR && act <= DS_ACT_SOFTPLUS, "ds_bn_bias_act_forward: unknown activation id %d", act);
  cudaStream_t st = (cudaStream_t)stream;
  bn_finalize_kernel<<<(unsigned)((F + 127) / 128), 128, 0, st>>>(F, sums, count, count_dev, eps, momentum, training,
                                                                  moving_mean, moving_var, bias, mean_rstd, scale_shift);
  DS_LAUNCHED();
  const int64_t n = B * M * F;
  if (F % 4 == 0) bn_apply_kernel<4><<<stream_grid(n / 4), 256, 0, st>>>(n / 4, F / 4, z, scale_shift, act, y);
  else bn_apply_kernel<1><<<stream_grid(n), 256, 0, st>>>(n, F, z, scale_shift, act, y);
  DS_LAUNCHED();
  return 0;
}

extern "C" int ds_bn_backward_stats(int64_t B, int64_t M, int64_t F, int64_t r0, int64_t r1, const float* z,
                                    const float* y, const float* dy, const float* mean_rstd, int32_t act, double* sums,
                                    double* workspace, void* stream) {
  using namespace ds;
  DS_CHECK(B > 0 && M > 0 && F > 0 && r0 >= 0 && r0 < r1 && r1 <= M, "ds_bn_backward_stats: bad sizes");
  DS_CHECK(z && dy && mean_rstd && sums && workspace && (act == DS_ACT_LINEAR || y), "ds_bn_backward_stats: NULL pointer");
  return bn_sums<1>(B, M, F, r0, r1, z, y, dy, mean_rstd, act, sums, workspace, (cudaStream_t)stream);
}

extern "C" int ds_bn_backward_apply(int64_t B, int64_t M, int64_t F, int64_t r0, int64_t r1, const float* z,
                                    const float* y, const float* dy, const float* mean_rstd, const double* sums,
                                    double count, const double* count_dev, int32_t act, int32_t training, float* dz,
                                    float* dbias, void* stream) {
  using namespace ds;
  DS_CHECK(B > 0 && M > 0 && F > 0 && r0 >= 0 && r0 < r1 && r1 <= M && (count > 0 || count_dev != nullptr),
           "ds_bn_backward_apply: bad sizes");
  DS_CHECK(z && dy && mean_rstd && sums && dz && (act == DS_ACT_LINEAR || y), "ds_bn_backward_apply: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  bn_backward_apply_kernel<<<stream_grid(B * M * F), 256, 0, st>>>(B, M, F, r0, r1, z, y, dy, mean_rstd, sums, count,
                                                                  count_dev, act, training, dz);
  DS_LAUNCHED();
  if (dbias != nullptr) {
    bn_dbias_kernel<<<(unsigned)((F + 127) / 128), 128, 0, st>>>(F, sums, dbias);
    DS_LAUNCHED();
  }
  return 0;
}
