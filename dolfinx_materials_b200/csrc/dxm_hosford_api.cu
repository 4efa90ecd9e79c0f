// Launch side of the Hosford behaviour (its own translation unit: the local-solve kernels are large and are
// instantiated once per supported compile-time exponent).
#define DXM_HOSFORD_KERNELS
#include "dxm_internal.cuh"
#include "dxm_hosford.cuh"

namespace dxm {

// sup over all stress states of sigma_eq(Hosford, a) / seq(von Mises) = (2^(a-1) + 1)^(1/a) / sqrt(3), reached in
// pure shear (1 for a = 2 and a = 4, -> 2/sqrt(3) = Tresca as a -> inf); tests/test_oracle_hosford.py scans it.
double hosford_bound(int a) {
  return std::pow(std::pow(2.0, (double)(a - 1)) + 1.0, 1.0 / (double)a) / std::sqrt(3.0) * (1.0 + 1e-9);
}

namespace {
template <int AT, bool VOCE, int MINB>
int launch_at2(const SmallStrainArgs& a, const HosLaunch& cfg, int* launches) {
  if (cfg.tiled) {
    const int64_t ntile = (a.count + kHosTile - 1) / kHosTile;
    dxm_hosford_tiled_kernel<AT, VOCE, MINB><<<(unsigned)(ntile < 1 ? 1 : ntile), kHosBlock, 0, cfg.stream>>>(a);
    ++*launches;
    CK(cudaGetLastError());
    return 0;
  }
  // small batches: 32-point CTAs of one tile each over all SMs (latency-bound; see launch_small_strain in dxm_api.cu)
  const bool small = a.count <= (int64_t)cfg.num_sms * 128;
  const int block = small ? 32 : kHosBlock;
  const int64_t ntile = (a.count + block - 1) / block;
  int64_t grid = small ? ntile : (ntile + cfg.tiles_per_cta - 1) / cfg.tiles_per_cta;
  if (grid < 1) grid = 1;
  dxm_hosford_kernel<AT, VOCE, MINB><<<(unsigned)grid, block, 0, cfg.stream>>>(a);
  ++*launches;
  CK(cudaGetLastError());
  return 0;
}

template <int AT, bool VOCE>
int launch_at(const SmallStrainArgs& a, const HosLaunch& cfg, int* launches) {
  // Registers: the fused kernel gains 6 % at 4 resident CTAs per SM (128 registers) now that only the Newton loop's own
  // state crosses the loop; the tiled kernel, whose streaming phase shares the allocation, loses 4-13 % there and stays
  // at 3 CTAs (168 registers) -- profiles/r02c_hosford_ab_minb_fused_warpqueue.json; 5 CTAs (96 registers, 0.9 KB of
  // spills) lose 9 % again (profiles/r02h_hosford_ab_minb5.json).  DXM_HOS_MINB=3|4 forces either.
  const int minb = cfg.minb ? cfg.minb : (cfg.tiled ? 3 : 4);
  if (minb == 3) return launch_at2<AT, VOCE, 3>(a, cfg, launches);
#ifdef DXM_HOS_TRY_MINB  // experiment builds (DXM_VARIANT): another register target in place of the 128-register one
  return launch_at2<AT, VOCE, DXM_HOS_TRY_MINB>(a, cfg, launches);
#else
  return launch_at2<AT, VOCE, 4>(a, cfg, launches);
#endif
}
}  // namespace

// exponents with an unrolled instantiation: 6 and 8 (the usual bcc / fcc fits) and 10 (the reference demo); any other
// even exponent runs the generic loops -- same operation order, same bits
int launch_hosford(const SmallStrainArgs& a, const HosLaunch& cfg, int* launches) {
  if (cfg.voce) {
    switch (a.hos_a) {
      case 6: return launch_at<6, true>(a, cfg, launches);
      case 8: return launch_at<8, true>(a, cfg, launches);
      case 10: return launch_at<10, true>(a, cfg, launches);
      default: return launch_at<0, true>(a, cfg, launches);
    }
  }
  switch (a.hos_a) {
    case 6: return launch_at<6, false>(a, cfg, launches);
    case 8: return launch_at<8, false>(a, cfg, launches);
    case 10: return launch_at<10, false>(a, cfg, launches);
    default: return launch_at<0, false>(a, cfg, launches);
  }
}

}  // namespace dxm
