// Runs the fused lattice convolution kernel of csrc/ds_lattice_conv2.cu (warp-specialised: compute warps, UMMA issuer,
// gather warps; mbarriers, cp.async, tcgen05 + tensor memory) ON THE HOST through tests/emul, on a problem prepared by
// tests/test_emul_cpu.py (tile tables of the lattice plan, input, weights), and writes y (and optionally the basis).
//   g++ -std=c++20 -O1 -pthread -DDS_EMULATE -I tests/emul tests/emul/emul_conv2.cpp -o emul_conv2
//   ./emul_conv2 <dir>      reads <dir>/meta.txt pix.bin w.bin x.bin W.bin [bias.bin], writes <dir>/y.bin [u1..u4.bin]
#include "cuda_runtime.h"

#include <cstdio>
#include <fstream>
#include <string>

#include "../../deepsphere-cosmo-tf2_b200/csrc/ds_lattice_conv2.cu"

namespace ds {
std::atomic<int64_t> g_launches{0};
int fail(const char*, ...) { return 1; }
}  // namespace ds

template <class T>
static T* load(const std::string& path, size_t n) {
  T* p = static_cast<T*>(std::aligned_alloc(64, ((n * sizeof(T) + 63) / 64) * 64 + 64));
  std::ifstream f(path, std::ios::binary);
  if (!f.read(reinterpret_cast<char*>(p), (std::streamsize)(n * sizeof(T)))) {
    std::fprintf(stderr, "cannot read %zu items from %s\n", n, path.c_str());
    std::exit(2);
  }
  return p;
}
static void store(const std::string& path, const float* p, size_t n) {
  std::ofstream f(path, std::ios::binary);
  f.write(reinterpret_cast<const char*>(p), (std::streamsize)(n * 4));
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const std::string dir = argv[1];
  long n_tiles, B, M, F, N, nsteps, cheb, act, has_bias, grid, b_split, want_basis, bwd = 0;
  {
    std::ifstream m(dir + "/meta.txt");
    m >> n_tiles >> B >> M >> F >> N >> nsteps >> cheb >> act >> has_bias >> grid >> b_split >> want_basis >> bwd;
  }
  using namespace ds;
  const int K = (int)nsteps + 1, n_chunks = (int)(F / C2_FC);
  int32_t* pix = load<int32_t>(dir + "/pix.bin", (size_t)n_tiles * C2_P);
  float* w = load<float>(dir + "/w.bin", (size_t)n_tiles * C2_P * 9);
  float* x = load<float>(dir + "/x.bin", (size_t)(B * M * F));
  float* W = load<float>(dir + "/W.bin", (size_t)(F * K * N));
  float* bias = has_bias ? load<float>(dir + "/bias.bin", (size_t)N) : nullptr;
  const size_t ny = (size_t)(B * M * N), nu = (size_t)(B * M * F);
  float* y = static_cast<float*>(std::aligned_alloc(64, ny * 4 + 64));
  for (size_t i = 0; i < ny; ++i) y[i] = NAN;
  float* u[C2_H] = {nullptr, nullptr, nullptr, nullptr};
  for (int s = 0; s < nsteps && want_basis; ++s) {
    u[s] = static_cast<float*>(std::aligned_alloc(64, nu * 4 + 64));
    for (size_t i = 0; i < nu; ++i) u[s][i] = NAN;
  }
  // weight images (conv2_prep_b_kernel).  Forward launch: B_k(f, n) = W[(f*K + k)*N + n] (ds_graph_conv_forward);
  // backward-data launch on dz (F = the layer's Fout, N = its Fin): B_k(o, f) = kernel[(f*K + k)*F + o]
  // (ds_graph_conv_backward: strides 1, F, K*F)
  const size_t img_elems = (size_t)n_chunks * K * N * C2_FC;
  float* img = static_cast<float*>(std::aligned_alloc(64, img_elems * 4 + 64));
  const int64_t s_f = bwd ? 1 : (int64_t)K * N, s_k = bwd ? F : N, s_n = bwd ? (int64_t)K * F : 1;
  emul::launch(2, 256, [&] { conv2_prep_b_kernel(W, s_f, s_k, s_n, n_chunks, K, (int)N, img); });

  Conv2Args a;
  a.n_tiles = (int)n_tiles; a.pix = pix; a.w = w;
  a.B = B; a.M = M; a.F = (int)F; a.N = (int)N; a.b_split = (int)b_split; a.nsteps = (int)nsteps;
  a.wscale = cheb ? 2.f : 1.f;
  a.wdiag = 0.f;  // C2_CDIAG builds: the diagonal of L~ is one scalar (ds_plan_attach_lattice checks it on the host)
  for (size_t i = 0; i < (size_t)n_tiles * C2_P; ++i)
    if (pix[i] >= 0) { a.wdiag = a.wscale * w[i * 9 + 8]; break; }
  a.dbg = nullptr; a.sleep_mma = 0; a.sleep_ld = 0;
  a.in0 = x;
  for (int s = 0; s < C2_H; ++s) a.out[s] = u[s];
  a.b_img = img; a.bias = bias; a.act = (int)act; a.y = y;
  const size_t smem = conv2_smem_bytes((int)N, (int)nsteps);
  std::printf("emul_conv2: %ld tiles, B %ld, F %ld -> N %ld, %ld hops, grid %ld x %d threads, %zu bytes smem\n", n_tiles, B, F, N,
              nsteps, grid, C2_THREADS, smem);
  if (cheb) emul::launch((unsigned)grid, C2_THREADS, [&] { lattice_conv2_kernel<true>(a); }, smem);
  else emul::launch((unsigned)grid, C2_THREADS, [&] { lattice_conv2_kernel<false>(a); }, smem);
  store(dir + "/y.bin", y, ny);
  for (int s = 0; s < nsteps && want_basis; ++s) store(dir + "/u" + std::to_string(s + 1) + ".bin", u[s], nu);
  std::printf("emul_conv2: done\n");
  return 0;
}
