// The library's own NCCL communicator: one rank per GPU, used for exactly one thing -- the in-stream all-gather of the
// per-call statistics record (64 bytes per rank: failed / plastic point counts, iteration and residual maxima) that
// follows the update kernel on the handle's stream when a handle has global statistics switched on
// (dxm_use_global_stats).  The constitutive update itself has no exchange step: Gauss points are
// independent, state never leaves its GPU.
//
// On one node the all-gather does not go through NCCL at all: every rank maps every other rank's exchange buffer
// (cudaIpc handles over NVLink / NVSwitch) and the update kernel's publishing CTA stores its record into all of them,
// waits for the peers' records in its own buffer and folds them -- the collective is part of the kernel's epilogue
// (csrc/dxm_canon.cuh::finish_record).  NCCL stays as the fallback when peer mapping is not available.
//
// NCCL is not linked: the functions are looked up at run time in the libnccl.so.2 already loaded into the process (the
// one PyTorch ships, when the caller is a torch.distributed program) or found by the dynamic loader (DXM_NCCL_LIB
// overrides the name).  The unique id travels between the ranks by whatever the caller has -- torch.distributed
// broadcast in dolfinx_materials_b200.distributed.init_stats_comm, MPI_Bcast under dolfinx.
#include <dlfcn.h>

#include "dxm_internal.cuh"

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;
constexpr int kNcclUint8 = 1;  // ncclDataType_t::ncclUint8

struct Nccl {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1, device = -1;
  // peer-memory exchange
  dxm::StatRecord* xchg_local = nullptr;            // [kXchgSlots][2][nranks]
  unsigned generation = 0;                          // bumped by every dxm_comm_init: slots belong to one communicator
  dxm::StatRecord* xchg_peer[dxm::kXchgMaxRanks] = {};  // mapped peers (own entry = xchg_local)
  dxm::StatXchg* d_desc = nullptr;                  // device copy of the descriptor the kernels read
  bool p2p = false;
  int next_slot = 0;
} g;

int load_nccl() {
  if (g.lib) return 0;
  const char* name = std::getenv("DXM_NCCL_LIB");
  void* lib = dlopen(name ? name : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return fail(std::string("dxm_comm: cannot load NCCL (") + (dlerror() ? dlerror() : "dlopen failed") + ")");
  g.GetUniqueId = (decltype(g.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
  g.CommInitRank = (decltype(g.CommInitRank))dlsym(lib, "ncclCommInitRank");
  g.AllGather = (decltype(g.AllGather))dlsym(lib, "ncclAllGather");
  g.CommDestroy = (decltype(g.CommDestroy))dlsym(lib, "ncclCommDestroy");
  g.GetErrorString = (decltype(g.GetErrorString))dlsym(lib, "ncclGetErrorString");
  if (!g.GetUniqueId || !g.CommInitRank || !g.AllGather || !g.CommDestroy || !g.GetErrorString) {
    dlclose(lib);
    return fail("dxm_comm: the NCCL library lacks an expected symbol");
  }
  g.lib = lib;
  return 0;
}

int nccl_check(ncclResult_t r, const char* what) {
  if (r == 0) return 0;
  return fail(std::string("dxm_comm: ") + what + ": " + g.GetErrorString(r));
}
}  // namespace

namespace dxm_comm {
const dxm::StatXchg* xchg() { return g.p2p ? g.d_desc : nullptr; }
int xchg_slot() { return g.next_slot < dxm::kXchgSlots ? g.next_slot++ : -1; }
unsigned generation() { return g.generation; }
int size() { return g.comm ? g.nranks : 1; }
int rank() { return g.comm ? g.rank : 0; }

int all_gather(const void* send, void* recv, size_t bytes_per_rank, cudaStream_t stream) {
  if (!g.comm) return fail("dxm_comm: no communicator (dxm_comm_init)");
  return nccl_check(g.AllGather(send, recv, bytes_per_rank, kNcclUint8, g.comm, stream), "ncclAllGather");
}
}  // namespace dxm_comm

extern "C" {

int dxm_comm_unique_id(void* id128) {
  if (!id128) return fail("dxm_comm_unique_id: NULL argument");
  if (load_nccl()) return -1;
  ncclUniqueId id;
  if (nccl_check(g.GetUniqueId(&id), "ncclGetUniqueId")) return -1;
  std::memcpy(id128, id.internal, sizeof(id.internal));
  return 0;
}

int dxm_comm_init(const void* id128, int rank, int nranks, int device) {
  if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail("dxm_comm_init: bad argument");
  if (g.comm) return fail("dxm_comm_init: a communicator already exists (dxm_comm_destroy first)");
  if (load_nccl()) return -1;
  CK(cudaSetDevice(device));
  ncclUniqueId id;
  std::memcpy(id.internal, id128, sizeof(id.internal));
  ncclComm_t c = nullptr;
  if (nccl_check(g.CommInitRank(&c, nranks, id, rank), "ncclCommInitRank")) return -1;
  g.comm = c;
  ++g.generation;
  g.rank = rank;
  g.nranks = nranks;
  g.device = device;
  return 0;
}

// Peer-memory exchange, step 1: allocate this rank's buffer and hand out its 64-byte cudaIpc handle.
int dxm_comm_p2p_handle(void* handle64) {
  if (!handle64) return fail("dxm_comm_p2p_handle: NULL argument");
  if (!g.comm) return fail("dxm_comm_p2p_handle: no communicator (dxm_comm_init)");
  if (g.nranks > dxm::kXchgMaxRanks) return fail("dxm_comm_p2p_handle: too many ranks for the peer-memory exchange");
  CK(cudaSetDevice(g.device));
  const size_t bytes = sizeof(dxm::StatRecord) * dxm::kXchgSlots * 2 * g.nranks;
  if (!g.xchg_local) {
    CK(cudaMalloc((void**)&g.xchg_local, bytes));
    CK(cudaMemset(g.xchg_local, 0, bytes));
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t hd;
  CK(cudaIpcGetMemHandle(&hd, g.xchg_local));
  std::memcpy(handle64, &hd, 64);
  return 0;
}

// Step 2 (after the caller gathered the handles of all ranks, rank order): map the peers.  On failure the exchange stays
// off and global statistics go through NCCL.
int dxm_comm_p2p_connect(const void* handles) {
  if (!handles || !g.comm || !g.xchg_local) return fail("dxm_comm_p2p_connect: call dxm_comm_p2p_handle first");
  CK(cudaSetDevice(g.device));
  dxm::StatXchg desc{};
  desc.nranks = g.nranks;
  desc.rank = g.rank;
  for (int r = 0; r < g.nranks; ++r) {
    if (r == g.rank) {
      g.xchg_peer[r] = g.xchg_local;
    } else {
      cudaIpcMemHandle_t hd;
      std::memcpy(&hd, (const char*)handles + 64 * r, 64);
      void* p = nullptr;
      const cudaError_t e = cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        cudaGetLastError();
        for (int q = 0; q < r; ++q)
          if (q != g.rank && g.xchg_peer[q]) cudaIpcCloseMemHandle(g.xchg_peer[q]);
        for (int q = 0; q < dxm::kXchgMaxRanks; ++q) g.xchg_peer[q] = nullptr;
        return fail(std::string("dxm_comm_p2p_connect: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
      }
      g.xchg_peer[r] = (dxm::StatRecord*)p;
    }
    desc.peer[r] = g.xchg_peer[r];
  }
  if (!g.d_desc) CK(cudaMalloc((void**)&g.d_desc, sizeof(dxm::StatXchg)));
  CK(cudaMemcpy(g.d_desc, &desc, sizeof(desc), cudaMemcpyHostToDevice));
  g.p2p = true;
  return 0;
}

int dxm_comm_p2p_enabled(void) { return g.p2p ? 1 : 0; }

// every rank must end up with the same answer: the caller switches the exchange off everywhere if any rank failed
int dxm_comm_p2p_disable(void) {
  g.p2p = false;
  return 0;
}

int dxm_comm_size(void) { return dxm_comm::size(); }
int dxm_comm_rank(void) { return dxm_comm::rank(); }

int dxm_comm_destroy(void) {
  if (!g.comm) return 0;
  cudaSetDevice(g.device);
  cudaDeviceSynchronize();
  for (int q = 0; q < dxm::kXchgMaxRanks; ++q) {
    if (q != g.rank && g.xchg_peer[q]) cudaIpcCloseMemHandle(g.xchg_peer[q]);
    g.xchg_peer[q] = nullptr;
  }
  cudaFree(g.xchg_local);
  cudaFree(g.d_desc);
  g.xchg_local = nullptr;
  g.d_desc = nullptr;
  g.p2p = false;
  g.next_slot = 0;
  const ncclResult_t r = g.CommDestroy(g.comm);
  g.comm = nullptr;
  g.nranks = 1;
  g.rank = 0;
  return nccl_check(r, "ncclCommDestroy");
}

}  // extern "C"
