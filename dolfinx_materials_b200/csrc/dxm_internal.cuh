// Shared internals of libdxm_cuda.so's translation units (not part of the ABI): error plumbing, launch accounting
// and the material handle.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/dxm.h"
#include "dxm_canon.cuh"

namespace dxm_detail {
extern thread_local std::string g_err;        // text behind dxm_last_error(), per calling thread
extern std::atomic<long long> g_launches;     // kernels launched by the library (dxm_launch_count)
inline int fail(const std::string& msg) {
  g_err = msg;
  return -1;
}
}  // namespace dxm_detail
using dxm_detail::fail;
using dxm_detail::g_err;
using dxm_detail::g_launches;

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + \
                  std::to_string(__LINE__) + ")");                                            \
    }                                                                                         \
  } while (0)

#define LAUNCH_CHECK()            \
  do {                            \
    g_launches.fetch_add(1);      \
    CK(cudaGetLastError());       \
  } while (0)

constexpr int kNProp = 6;

struct Field {
  const char* name;
  int row;  // first SoA row inside a generation block
  int dim;
};

struct dxm_handle {
  int behaviour = 0, device = 0;
  int64_t n = 0, ld = 0;
  int ngrad = 0, nflux = 0, nisv = 0, nrows = 0, nct = 0;
  int nct_store = 0;  // resident tangent rows: nct, or 21 for the packed symmetric 6x6 of the small-strain behaviours
  std::vector<Field> fields;
  double* gen[2] = {nullptr, nullptr};  // device SoA blocks [nrows][ld]
  int i0 = 0;                           // gen[i0] is s0, gen[1-i0] is s1
  bool s1_valid = false;                // false => s1 reads alias s0 (after update/revert)
  double* ct = nullptr;                 // [nct_store][ld]
  // properties
  double uni[kNProp] = {0, 0, 0, 0, 0, 0};
  bool set[kNProp] = {false, false, false, false, false, false};
  bool perpoint = false;
  double* pp = nullptr;  // [kNProp][ld]
  double* table = nullptr;  // DXM_J2_TABLE: device [3][ntab] = p_k, sig_k, slope_k
  int ntab = 0;
  int hos_a = 10;  // DXM_HOSFORD_LINEAR: exponent (the demo's value by default)
  // statistics (dxm_canon.cuh): accumulated, folded and published by the update kernel itself
  dxm::StatBlock* d_statblk = nullptr;   // 32 accumulation slots + ticket, cleared by the kernel that folds them
  dxm::StatRecord* d_rec = nullptr;      // this rank's record on the device (source of the in-stream all-gather)
  dxm::StatRecord* d_gather = nullptr;   // [nranks] records after the all-gather
  dxm::StatRecord* h_rec = nullptr;      // page-locked, mapped: the published record, polled by the host
  unsigned long long seq = 0;            // sequence number of the last call launched on this handle
  bool finalize_launched = false;        // the call's last launch (the one that publishes) has been enqueued
  bool global_stats = false;             // reduce over the ranks of the library's communicator (dxm_comm_init)
  int xslot = -1;                        // >= 0: this handle's slot in the peer-memory exchange buffers (else: NCCL)
  unsigned xgen = 0;                     // communicator generation the slot was drawn from
  int timing = -1;                       // kernel_ms events: -1 auto (batches >= 262144 points), 0 never, 1 always
  dxm_stats last{};
  bool stats_pending = false;
  // streams / events
  cudaStream_t stream = nullptr, own_stream = nullptr, s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_in_free[2] = {nullptr, nullptr},
              ev_packed[2] = {nullptr, nullptr}, ev_out_free[2] = {nullptr, nullptr};
  std::vector<cudaEvent_t> ev_k;  // kernel timing event pairs
  int n_ev_used = 0;
  // staging (device AoS), allocated lazily
  int64_t chunk = 0;
  double* d_in[2] = {nullptr, nullptr};
  double* d_out[2] = {nullptr, nullptr};
  double* h_small = nullptr;  // mapped page-locked staging of the small-batch host path (dxm_api.cu: kSmallHostPoints)
  // host mirror of the packed tangent (dxm_host_mirror.hpp): page-locked ring of packed (chunk, 21) blocks
  static constexpr int kRing = 3;
  double* h_ctp[kRing] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_ct[kRing] = {nullptr, nullptr, nullptr};
  // diagnostics
  bool diag = false;
  uint8_t *d_flag = nullptr, *d_fail = nullptr;
  int32_t* d_iter = nullptr;
  double* d_resid = nullptr;
  int num_sms = 148;
  int ppt = 1;
  int minb = 0;
  int vote = 1;
  int compact = -1;  // DXM_COMPACT: 0 never, 1 always, unset = auto (FeFp only, by the last plastic fraction)
  int64_t prev_plastic = 0, prev_points = 0;
  std::atomic<int> refs{1};
};

// the library's NCCL communicator (dxm_comm.cu): in-stream all-gather of the statistics records, nothing else
namespace dxm_comm {
int size();
int rank();
int all_gather(const void* send, void* recv, size_t bytes_per_rank, cudaStream_t stream);
const dxm::StatXchg* xchg();  // device descriptor of the peer-memory exchange, or null (then: NCCL)
int xchg_slot();              // next free exchange slot (same order on every rank), -1 when exhausted
unsigned generation();        // which dxm_comm_init the slots belong to
}  // namespace dxm_comm

inline int set_device(const dxm_handle* h) {
  CK(cudaSetDevice(h->device));
  return 0;
}
