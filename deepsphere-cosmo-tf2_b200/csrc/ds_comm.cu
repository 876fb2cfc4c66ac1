// ds_comm_*: thin wrappers over NCCL for the two exchanges of the path so that a binder without a
// collective library of its own (the TensorFlow shim of INTEGRATION.md) can run the multi-GPU path through this C-ABI alone:
//   - the all-to-all of halo rows of a sphere-partitioned graph convolution (grouped ncclSend / ncclRecv to the peers,
//     all equidistant through NVSwitch), with the pack / assemble / reduce kernels of ds_halo.cu around it in ONE call;
//   - the all-reduce of weight gradients and of the 2F + 1 BatchNorm sums.
// NCCL is loaded at run time (dlopen), never linked: the library loads and every single-GPU entry works on a machine
// without NCCL; ds_comm_* then fail with a message.  The Python host of this repository uses torch.distributed's NCCL
// communicator for the same exchanges (deepsphere/partition.py) and calls only the ds_halo_* kernels.
#include <dlfcn.h>

#include <mutex>
#include <vector>

#include "ds_common.cuh"

namespace ds {
namespace {

// the handful of NCCL declarations needed (nccl.h: NCCL_UNIQUE_ID_BYTES 128, ncclSum 0, ncclFloat32 7, ncclFloat64 8)
typedef struct { char internal[128]; } NcclUniqueId;
typedef void* NcclComm;
struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  void* handle = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

int load_nccl(const char* path) {
  std::lock_guard<std::mutex> lock(g_nccl_mu);
  if (g_nccl.handle != nullptr) return 0;
  const char* candidates[] = {path, getenv("DEEPSPHERE_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* c : candidates) {
    if (c == nullptr || *c == 0) continue;
    h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (h != nullptr) break;
  }
  if (h == nullptr) return fail("ds_comm: cannot load NCCL (%s); pass its path to ds_comm_load_nccl or set DEEPSPHERE_NCCL_LIB", dlerror());
#define DS_SYM(field, name)                                                              \
  *reinterpret_cast<void**>(&g_nccl.field) = dlsym(h, name);                             \
  if (g_nccl.field == nullptr) return fail("ds_comm: symbol %s not found in the NCCL library", name)
  DS_SYM(GetUniqueId, "ncclGetUniqueId");
  DS_SYM(CommInitRank, "ncclCommInitRank");
  DS_SYM(CommDestroy, "ncclCommDestroy");
  DS_SYM(AllReduce, "ncclAllReduce");
  DS_SYM(Send, "ncclSend");
  DS_SYM(Recv, "ncclRecv");
  DS_SYM(GroupStart, "ncclGroupStart");
  DS_SYM(GroupEnd, "ncclGroupEnd");
  DS_SYM(GetErrorString, "ncclGetErrorString");
#undef DS_SYM
  g_nccl.handle = h;
  return 0;
}

#define DS_NCCL(expr)                                                                                   \
  do {                                                                                                  \
    const int _r = (expr);                                                                              \
    if (_r != 0) return ds::fail("%s failed: %s", #expr, ds::g_nccl.GetErrorString ? ds::g_nccl.GetErrorString(_r) : "?"); \
  } while (0)

}  // namespace
}  // namespace ds

struct ds_comm {
  ds::NcclComm comm = nullptr;
  int rank = 0, world = 1;
};

extern "C" {

int ds_comm_load_nccl(const char* path) { return ds::load_nccl(path); }

int ds_comm_unique_id(char* id128) {
  using namespace ds;
  DS_CHECK(id128 != nullptr, "ds_comm_unique_id: NULL buffer");
  DS_TRY(load_nccl(nullptr));
  NcclUniqueId id;
  DS_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, id.internal, 128);
  return 0;
}

int ds_comm_create(int32_t world, int32_t rank, const char* id128, ds_comm_t** out) {
  using namespace ds;
  DS_CHECK(out != nullptr && id128 != nullptr && world >= 1 && rank >= 0 && rank < world, "ds_comm_create: bad argument");
  DS_TRY(load_nccl(nullptr));
  NcclUniqueId id;
  memcpy(id.internal, id128, 128);
  ds_comm* c = new ds_comm();
  c->rank = rank;
  c->world = world;
  const int r = g_nccl.CommInitRank(&c->comm, world, id, rank);
  if (r != 0) {
    delete c;
    return fail("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
  }
  *out = c;
  return 0;
}

int ds_comm_destroy(ds_comm_t* c) {
  if (c == nullptr) return 0;
  if (c->comm != nullptr && ds::g_nccl.CommDestroy) ds::g_nccl.CommDestroy(c->comm);
  delete c;
  return 0;
}

int ds_comm_allreduce_sum(ds_comm_t* c, void* buf, int64_t n, int32_t is_double, void* stream) {
  using namespace ds;
  DS_CHECK(c != nullptr && buf != nullptr && n >= 0, "ds_comm_allreduce_sum: bad argument");
  if (n == 0 || c->world == 1) return 0;
  DS_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, is_double ? 8 : 7, 0, c->comm, (cudaStream_t)stream));
  return 0;
}

// send[ sum_q send_counts[q] ] -> recv[ sum_q recv_counts[q] ] floats, peer blocks in rank order
int ds_comm_alltoallv(ds_comm_t* c, const float* send, const int64_t* send_counts, float* recv,
                      const int64_t* recv_counts, void* stream) {
  using namespace ds;
  DS_CHECK(c != nullptr && send_counts != nullptr && recv_counts != nullptr, "ds_comm_alltoallv: bad argument");
  if (c->world == 1) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  DS_NCCL(g_nccl.GroupStart());
  int64_t so = 0, ro = 0;
  for (int q = 0; q < c->world; ++q) {
    if (q != c->rank && send_counts[q] > 0) DS_NCCL(g_nccl.Send(send + so, (size_t)send_counts[q], 7, q, c->comm, st));
    if (q != c->rank && recv_counts[q] > 0) DS_NCCL(g_nccl.Recv(recv + ro, (size_t)recv_counts[q], 7, q, c->comm, st));
    so += send_counts[q];
    ro += recv_counts[q];
  }
  DS_NCCL(g_nccl.GroupEnd());
  return 0;
}

// One call = the whole forward halo exchange of a partitioned layer: pack (all peers) -> all-to-all -> assemble.
// send_rows / recv_pos: the concatenated row lists of ds_halo_pack / ds_halo_assemble; *_rows_per_peer [world]: how many
// of them belong to each peer; workspace: (n_send + n_recv) * B * F floats.
int ds_halo_exchange(ds_comm_t* c, int64_t B, int64_t n_own, int64_t n_ext, int64_t own_start, int64_t F,
                     const int32_t* send_rows, const int64_t* send_rows_per_peer, const int32_t* recv_pos,
                     const int64_t* recv_rows_per_peer, const float* x_own, float* x_ext, float* workspace, void* stream) {
  using namespace ds;
  DS_CHECK(c != nullptr && send_rows_per_peer != nullptr && recv_rows_per_peer != nullptr, "ds_halo_exchange: bad argument");
  int64_t n_send = 0, n_recv = 0;
  std::vector<int64_t> sc(c->world), rc(c->world);
  for (int q = 0; q < c->world; ++q) {
    n_send += send_rows_per_peer[q];
    n_recv += recv_rows_per_peer[q];
    sc[q] = send_rows_per_peer[q] * B * F;
    rc[q] = recv_rows_per_peer[q] * B * F;
  }
  DS_CHECK(workspace != nullptr || n_send + n_recv == 0, "ds_halo_exchange: workspace required");
  float* send = workspace;
  float* recv = workspace + n_send * B * F;
  DS_TRY(ds_halo_pack(B, n_own, F, n_send, send_rows, x_own, send, stream));
  DS_TRY(ds_comm_alltoallv(c, send, sc.data(), recv, rc.data(), stream));
  return ds_halo_assemble(B, n_own, n_ext, own_start, F, n_recv, recv_pos, x_own, recv, x_ext, stream);
}

// ... and the transposed exchange of the backward pass: pack the halo rows of g_ext -> all-to-all -> reduce into g_own.
int ds_halo_exchange_backward(ds_comm_t* c, int64_t B, int64_t n_own, int64_t n_ext, int64_t own_start, int64_t F,
                              const int64_t* send_rows_per_peer, const int32_t* recv_pos, const int64_t* recv_rows_per_peer,
                              const int32_t* row_slot_ptr, const int32_t* slots, const float* g_ext, float* g_own,
                              float* workspace, void* stream) {
  using namespace ds;
  DS_CHECK(c != nullptr && send_rows_per_peer != nullptr && recv_rows_per_peer != nullptr,
           "ds_halo_exchange_backward: bad argument");
  int64_t n_send = 0, n_recv = 0;
  std::vector<int64_t> sc(c->world), rc(c->world);
  for (int q = 0; q < c->world; ++q) {
    n_send += send_rows_per_peer[q];
    n_recv += recv_rows_per_peer[q];
    sc[q] = recv_rows_per_peer[q] * B * F;  // what came in goes back out
    rc[q] = send_rows_per_peer[q] * B * F;
  }
  DS_CHECK(workspace != nullptr || n_send + n_recv == 0, "ds_halo_exchange_backward: workspace required");
  float* back = workspace;
  float* got = workspace + n_recv * B * F;
  DS_TRY(ds_halo_pack(B, n_ext, F, n_recv, recv_pos, g_ext, back, stream));
  DS_TRY(ds_comm_alltoallv(c, back, sc.data(), got, rc.data(), stream));
  return ds_halo_reduce(B, n_own, n_ext, own_start, F, n_send > 0 ? row_slot_ptr : nullptr, slots, g_ext, got, g_own, stream);
}

}  // extern "C"
