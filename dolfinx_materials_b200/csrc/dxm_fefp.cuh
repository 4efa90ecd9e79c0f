// Finite-strain FeFp J2 plasticity kernel (placeholder until the kernel lands).
#pragma once
#include <atomic>
#include "dxm_canon.cuh"

namespace dxm {

struct FeFpArgs {
  const double* F;
  double* P;
  double* p;
  double* be;
  double* ct;
  const double* F_old;
  const double* p_old;
  const double* be_old;
  int64_t ld, start, count;
  bool perpoint;
  double E, mu, kappa, sig0, H, dsu, b;
  const double* pp[6];
  StatSlot* stats;
  uint8_t* d_flag;
  int32_t* d_iter;
  double* d_resid;
  uint8_t* d_fail;
};

int launch_fefp(const FeFpArgs& a, bool diag, int num_sms, cudaStream_t stream,
                std::atomic<long long>* launches);

}  // namespace dxm
