// Runs the kernels of csrc/ds_skinny.cu ON THE HOST (tests/emul/cuda_runtime.h) against naive loops.
//   g++ -std=c++20 -O1 -pthread -DDS_EMULATE -I tests/emul tests/emul/emul_skinny.cpp -o emul_skinny && ./emul_skinny
// Prints one "ok ..." line per case and exits 0, or "FAIL ..." and exits 1.
#include "cuda_runtime.h"

#include <algorithm>
#include <random>

#include "../../deepsphere-cosmo-tf2_b200/csrc/ds_skinny.cu"

namespace ds {  // symbols ds_common.cuh declares and the kernels never call
std::atomic<int64_t> g_launches{0};
int fail(const char*, ...) { return 1; }
}  // namespace ds

static float host_act(float v, int act) { return ds::act_apply(v, act); }

template <int KC>
static int run_case(int64_t R, int N, int act, unsigned grid, bool with_bias, bool want_dz, unsigned seed) {
  std::mt19937 rng(seed);
  std::normal_distribution<float> nd(0.f, 1.f);
  auto fill = [&](std::vector<float>& v) { for (auto& e : v) e = nd(rng); };
  // float4 accesses need 16-byte aligned bases: over-allocate and align by hand
  auto aligned = [](std::vector<float>& v) { return reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(v.data()) + 15) & ~uintptr_t(15)); };
  std::vector<float> Xs(R * KC + 4), Ws(KC * N + 4), bs(N), Ys(R * N + 4), dYs(R * N + 4), dZs(R * N + 4, -777.f);
  fill(Xs); fill(Ws); fill(bs); fill(dYs);
  float *X = aligned(Xs), *W = aligned(Ws), *Y = aligned(Ys), *dY = aligned(dYs), *dZ = aligned(dZs);
  // ---- forward
  emul::launch(grid, 256, [&] { ds::skinny_nn_kernel<KC>(R, N, X, W, with_bias ? bs.data() : nullptr, N, act, Y); });
  double err = 0;
  for (int64_t r = 0; r < R; ++r)
    for (int n = 0; n < N; ++n) {
      float acc = 0.f;
      for (int k = 0; k < KC; ++k) acc = std::fmaf(X[r * KC + k], W[k * N + n], acc);
      if (with_bias) acc += bs[n];
      err = std::max(err, (double)std::fabs(host_act(acc, act) - Y[r * N + n]));
    }
  if (err != 0) { std::printf("FAIL forward KC=%d R=%lld N=%d act=%d: max err %g (must be bit-exact)\n", KC, (long long)R, N, act, err); return 1; }
  // ---- backward sweep + final reduction
  const int nblk = (int)grid;
  std::vector<float> partial((size_t)nblk * (KC + 1) * N, -555.f), dw(KC * N, -1.f), db(N, -1.f);
  emul::launch(grid, 256, [&] { ds::skinny_bwd_kernel<KC>(R, N, X, Y, dY, act, want_dz ? dZ : nullptr, partial.data()); });
  emul::launch(1, 128, [&] { ds::skinny_bwd_final_kernel(N, KC, nblk, partial.data(), dw.data(), with_bias ? db.data() : nullptr); });
  std::vector<double> rdw(KC * N, 0.0), rdb(N, 0.0);
  double dz_err = 0, scale = 0;
  for (int64_t r = 0; r < R; ++r)
    for (int n = 0; n < N; ++n) {
      const float dz = act == DS_ACT_LINEAR ? dY[r * N + n] : dY[r * N + n] * ds::act_grad_from_y(Y[r * N + n], act);
      if (want_dz && act != DS_ACT_LINEAR) dz_err = std::max(dz_err, (double)std::fabs(dz - dZ[r * N + n]));
      rdb[n] += dz;
      for (int k = 0; k < KC; ++k) rdw[k * N + n] += (double)X[r * KC + k] * dz;
    }
  double e_dw = 0, e_db = 0;
  for (int i = 0; i < KC * N; ++i) { e_dw = std::max(e_dw, std::fabs(rdw[i] - dw[i])); scale = std::max(scale, std::fabs(rdw[i])); }
  for (int n = 0; n < N; ++n) e_db = std::max(e_db, std::fabs(rdb[n] - (with_bias ? db[n] : rdb[n])));
  const double tol = 2e-5 * std::max(scale, 1.0);
  if (dz_err != 0 || e_dw > tol || e_db > tol) {
    std::printf("FAIL backward KC=%d R=%lld N=%d act=%d grid=%u: dz %g dw %g db %g (tol %g)\n", KC, (long long)R, N, act, grid, dz_err, e_dw, e_db, tol);
    return 1;
  }
  if (!want_dz) for (int64_t i = 0; i < R * N; ++i) if (dZ[i] != -777.f) { std::printf("FAIL dz written although not requested\n"); return 1; }
  std::printf("ok KC=%d R=%lld N=%d act=%d grid=%u bias=%d dz=%d: fwd bit-exact, dw err %.2e db err %.2e\n", KC, (long long)R, N, act, grid, (int)with_bias, (int)want_dz, e_dw, e_db);
  return 0;
}

// act_backward_colsum_kernel: dz bit-exact, block partials [nblk][F] summing to the column sums
static int run_act_colsum(int64_t R, int F, int act, unsigned grid, unsigned seed) {
  std::mt19937 rng(seed);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::vector<float> Ys(R * F + 4), dYs(R * F + 4), dZs(R * F + 4, -777.f), partial((size_t)grid * F, -555.f);
  auto aligned = [](std::vector<float>& v) { return reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(v.data()) + 15) & ~uintptr_t(15)); };
  for (auto& e : Ys) e = nd(rng);
  for (auto& e : dYs) e = nd(rng);
  float *Y = aligned(Ys), *dY = aligned(dYs), *dZ = aligned(dZs);
  emul::launch(grid, 256, [&] { ds::act_backward_colsum_kernel(R, F, Y, dY, act, dZ, partial.data()); });
  double e_dz = 0, e_cs = 0, scale = 1;
  std::vector<double> cs(F, 0.0);
  for (int64_t r = 0; r < R; ++r)
    for (int f = 0; f < F; ++f) {
      const float dz = dY[r * F + f] * ds::act_grad_from_y(Y[r * F + f], act);
      e_dz = std::max(e_dz, (double)std::fabs(dz - dZ[r * F + f]));
      cs[f] += dz;
    }
  for (int f = 0; f < F; ++f) {
    double s = 0;
    for (unsigned b = 0; b < grid; ++b) s += partial[(size_t)b * F + f];
    e_cs = std::max(e_cs, std::fabs(s - cs[f]));
    scale = std::max(scale, std::fabs(cs[f]));
  }
  if (e_dz != 0 || e_cs > 2e-5 * scale) {
    std::printf("FAIL act_backward_colsum R=%lld F=%d act=%d grid=%u: dz %g colsum %g\n", (long long)R, F, act, grid, e_dz, e_cs);
    return 1;
  }
  std::printf("ok act_backward_colsum R=%lld F=%d act=%d grid=%u: dz bit-exact, colsum err %.2e\n", (long long)R, F, act, grid, e_cs);
  return 0;
}

int main() {
  int bad = 0;
  bad += run_act_colsum(1000, 32, DS_ACT_RELU, 3, 11);    // the HealpyGCNN layers: 32 / 64 channels
  bad += run_act_colsum(777, 64, DS_ACT_ELU, 4, 12);
  bad += run_act_colsum(5, 4, DS_ACT_TANH, 2, 13);        // cg = 1; second block has rows, later ones none
  bad += run_act_colsum(130, 256, DS_ACT_SIGMOID, 5, 14); // widest: 256 columns = 256 threads
  bad += run_act_colsum(64, 16, DS_ACT_SOFTPLUS, 7, 15);  // more blocks than row chunks
  bad += run_case<4>(1000, 16, DS_ACT_RELU, 3, true, false, 1);    // the HealpyGCNN head: 1 -> 16 channels, p = 1
  bad += run_case<4>(37, 4, DS_ACT_LINEAR, 2, false, false, 2);    // cg = 1, fewer rows than threads
  bad += run_case<8>(513, 8, DS_ACT_ELU, 1, true, true, 3);        // cg = 2, single block, dz requested
  bad += run_case<12>(300, 32, DS_ACT_TANH, 2, true, true, 4);     // cg = 8
  bad += run_case<16>(129, 64, DS_ACT_SIGMOID, 2, false, true, 5); // cg = 16, widest shape
  bad += run_case<16>(1, 64, DS_ACT_RELU, 1, true, false, 6);      // one row
  return bad ? 1 : 0;
}
