// Runs patch_conv_kernel of csrc/ds_patch.cu ON THE HOST (tests/emul/cuda_runtime.h) on a problem written by
// tests/test_emul_cpu.py: <dir>/meta.txt + binary arrays in, y.bin and u<s>.bin out (rows the kernel does not own stay NaN).
//   g++ -std=c++20 -O1 -pthread -DDS_EMULATE -I tests/emul tests/emul/emul_patch.cpp -o emul_patch && ./emul_patch <dir>
#include "cuda_runtime.h"

#include <atomic>
#include <fstream>
#include <limits>
#include <string>

#include "../../deepsphere-cosmo-tf2_b200/csrc/ds_patch.cu"

namespace ds {  // symbols ds_common.cuh declares and the kernel never calls
std::atomic<int64_t> g_launches{0};
int fail(const char*, ...) { return 1; }
}  // namespace ds

template <class T>
static std::vector<T> read_all(const std::string& path, size_t pad = 8) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) { std::fprintf(stderr, "cannot open %s\n", path.c_str()); std::exit(2); }
  const size_t bytes = (size_t)f.tellg();
  std::vector<T> v(bytes / sizeof(T) + pad);
  f.seekg(0);
  f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)bytes);
  return v;
}
static float* aligned16(std::vector<float>& v) {
  return reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(v.data()) + 15) & ~uintptr_t(15));
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const std::string d = argv[1];
  long long n_patches, B, M, F, N, nsteps, cheb, act, has_bias, want_basis, s_f, s_k, s_n, max_rows;
  {
    std::ifstream m(d + "/meta.txt");
    m >> n_patches >> B >> M >> F >> N >> nsteps >> cheb >> act >> has_bias >> want_basis >> s_f >> s_k >> s_n >> max_rows;
    if (!m) return 2;
  }
  auto row_ptr = read_all<int32_t>(d + "/row_ptr.bin");
  auto rows = read_all<int32_t>(d + "/rows.bin");
  auto ell_col = read_all<int32_t>(d + "/ell_col.bin");
  auto ell_val = read_all<float>(d + "/ell_val.bin");
  auto own_ptr = read_all<int32_t>(d + "/own_ptr.bin");
  auto own_local = read_all<int32_t>(d + "/own_local.bin");
  auto xs = read_all<float>(d + "/x.bin");
  auto W = read_all<float>(d + "/W.bin");
  auto bias = read_all<float>(d + "/bias.bin");
  // float4 accesses: 16-byte aligned copies
  std::vector<float> xa(B * M * F + 8);
  float* x = aligned16(xa);
  std::memcpy(x, xs.data(), sizeof(float) * B * M * F);
  const float nan = std::numeric_limits<float>::quiet_NaN();
  std::vector<float> ya(B * M * N + 8, nan);
  float* y = aligned16(ya);
  std::vector<std::vector<float>> us(nsteps);
  ds::PatchArgs a;
  a.row_ptr = row_ptr.data(); a.rows = rows.data(); a.ell_col = ell_col.data(); a.ell_val = ell_val.data();
  a.own_ptr = own_ptr.data(); a.own_local = own_local.data();
  a.B = B; a.M = M; a.F = (int)F; a.N = (int)N; a.nsteps = (int)nsteps; a.cheb = (int)cheb; a.act = (int)act;
  a.max_rows = (int)max_rows;
  a.in0 = x;
  for (int s = 0; s < ds::PATCH_MAX_STEPS; ++s) {
    a.out[s] = nullptr;
    if (want_basis && s < nsteps) {
      us[s].assign(B * M * F + 8, nan);
      a.out[s] = aligned16(us[s]);
    }
  }
  a.W = W.data(); a.s_f = s_f; a.s_k = s_k; a.s_n = s_n;
  a.bias = has_bias ? bias.data() : nullptr;
  a.y = y;
  const size_t smem = ds::patch_smem_bytes((int)max_rows, (int)F, (int)N);
  if (smem > 227 * 1024) { std::fprintf(stderr, "shared memory %zu too large\n", smem); return 2; }
  emul::launch((unsigned)(n_patches * B), ds::PATCH_THREADS, [&] { ds::patch_conv_kernel(a); }, smem);
  auto dump = [&](const std::string& name, const float* p, size_t n) {
    std::ofstream f(d + "/" + name, std::ios::binary);
    f.write(reinterpret_cast<const char*>(p), (std::streamsize)(n * sizeof(float)));
  };
  dump("y.bin", y, (size_t)(B * M * N));
  for (int s = 0; s < nsteps; ++s)
    if (a.out[s]) dump("u" + std::to_string(s + 1) + ".bin", a.out[s], (size_t)(B * M * F));
  std::printf("ok patch_conv patches=%lld B=%lld F=%lld N=%lld hops=%lld smem=%zu\n", n_patches, B, F, N, nsteps, smem);
  return 0;
}
