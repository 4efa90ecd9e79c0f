// CPU-side check of the CUDA kernel's per-point routine (tests/test_hosford_host.py): dxm::hosford_point is
// __host__ __device__, so the very code the kernel runs per Gauss point is executed here on the host, point by point,
// and compared bit for bit with the oracle -- without a GPU.  Test scaffolding only: nothing in the product calls this.
#include <cmath>

#include "../dolfinx_materials_b200/csrc/dxm_hosford.cuh"

namespace {
template <int AT, bool VOCE>
void run(int64_t n, const double* eps, const double* e_old, const double* s_old, const double* p_old,
         const double* ep_old, double E, double nu, double sig0, double H, double sigu, double b, int a, double bound, double* sig, double* p,
         double* epsp, double* ct, uint8_t* flag, int32_t* n_iter, double* resid, uint8_t* fail, int split,
         int64_t* n_candidates);
}

// same dispatch as launch_hosford (dxm_hosford_api.cu): unrolled instantiations for a = 6, 8, 10, generic loops else
extern "C" void hosford_host(int64_t n, const double* eps, const double* e_old, const double* s_old, const double* p_old,
                             const double* ep_old, double E, double nu, double sig0, double H, double sigu, double b, int a, double bound,
                             double* sig, double* p, double* epsp, double* ct, uint8_t* flag, int32_t* n_iter,
                             double* resid, uint8_t* fail, int split, int64_t* n_candidates, int force_generic) {
#define DXM_ARGS n, eps, e_old, s_old, p_old, ep_old, E, nu, sig0, H, sigu, b, a, bound, sig, p, epsp, ct, flag, n_iter, resid, fail, split, n_candidates
  const bool voce = sigu != sig0;  // the product keys on "sigu was set" (dxm_api.cu); equivalent for this harness
  if (voce) {
    if (force_generic) return run<0, true>(DXM_ARGS);
    switch (a) {
      case 6: return run<6, true>(DXM_ARGS);
      case 8: return run<8, true>(DXM_ARGS);
      case 10: return run<10, true>(DXM_ARGS);
      default: return run<0, true>(DXM_ARGS);
    }
  }
  if (force_generic) return run<0, false>(DXM_ARGS);
  switch (a) {
    case 6: return run<6, false>(DXM_ARGS);
    case 8: return run<8, false>(DXM_ARGS);
    case 10: return run<10, false>(DXM_ARGS);
    default: return run<0, false>(DXM_ARGS);
  }
}

namespace {
template <int AT, bool VOCE>
void run(int64_t n, const double* eps, const double* e_old, const double* s_old, const double* p_old,
         const double* ep_old, double E, double nu, double sig0, double H, double sigu, double b, int a, double bound, double* sig, double* p,
         double* epsp, double* ct, uint8_t* flag, int32_t* n_iter, double* resid, uint8_t* fail, int split,
         int64_t* n_candidates) {
  const double lam = E * nu / (1 + nu) / (1 - 2 * nu);
  const double mu = E / 2 / (1 + nu);
  double dsu = sigu - sig0;
  if (!std::isfinite(dsu)) dsu = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    double e1[6], e0[6], s0[6], ep0[6], so[6], epo[6], ct21[21], pn, rs;
    bool fl, fa;
    int it;
    for (int c = 0; c < 6; ++c) {
      e1[c] = eps[i * 6 + c];
      e0[c] = e_old[i * 6 + c];
      s0[c] = s_old[i * 6 + c];
      ep0[c] = ep_old[i * 6 + c];
    }
    // split != 0 replays the tiled kernel: phase A finishes the clearly elastic points and reports the candidates,
    // which the full routine then recomputes from scratch (dxm_hosford_tiled_kernel)
    bool heavy = true;
    if (split) heavy = dxm::hosford_point<true, 0, VOCE>(lam, mu, sig0, H, dsu, b, a, bound, e1, e0, s0, p_old[i], ep0, so, pn, epo, ct21, fl, it, rs, fa);
    if (heavy) {
      if (split) ++*n_candidates;
      dxm::hosford_point<false, AT, VOCE>(lam, mu, sig0, H, dsu, b, a, bound, e1, e0, s0, p_old[i], ep0, so, pn, epo, ct21, fl, it, rs, fa);
    }
    for (int c = 0; c < 6; ++c) {
      sig[i * 6 + c] = so[c];
      epsp[i * 6 + c] = epo[c];
    }
    p[i] = pn;
    for (int c = 0; c < 36; ++c) ct[i * 36 + c] = ct21[dxm::sym6_packed(c)];
    flag[i] = fl;
    n_iter[i] = it;
    resid[i] = rs;
    fail[i] = fa;
  }
}
}  // namespace
