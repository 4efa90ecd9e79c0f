// CPU-side check of the J2 and FeFp kernels' per-point routines (tests/test_point_host.py): dxm::j2_point,
// dxm::j2_tangent_entry, dxm::point_props and dxm::fefp_point are __host__ __device__, so the very code the kernels
// run per Gauss point is executed here on the host, point by point, and compared bit for bit with the oracle --
// without a GPU.  Test scaffolding only: nothing in the product calls this.  (Hosford: tests/hosford_host_check.cu.)
#include <cmath>
#include <vector>

#include "../dolfinx_materials_b200/csrc/dxm_fefp.cuh"
#include "../dolfinx_materials_b200/csrc/dxm_small_strain.cuh"

namespace {

// props: [6][n] per-point rows (E, nu, sig0, H, sigu, b) when perpoint, else 6 scalars
template <int HARD>
void run_j2(int64_t n, const double* eps, const double* e_old, const double* s_old, const double* p_old,
            const double* ep_old, const double* props, int perpoint, const double* table, int ntab, int vote,
            double* sig, double* p, double* epsp, double* ct, uint8_t* flag, int32_t* n_iter, double* resid,
            uint8_t* fail) {
  for (int64_t i = 0; i < n; ++i) {
    dxm::PointProps m;
    // uniform properties: the host side of launch_update (dxm_api.cu) derives the constants once, same expressions
    if (perpoint)
      dxm::point_props(props[0 * n + i], props[1 * n + i], props[2 * n + i], props[3 * n + i], props[4 * n + i],
                       props[5 * n + i], m);
    else
      dxm::point_props(props[0], props[1], props[2], props[3], props[4], props[5], m);
    m.tp = table;
    m.ts = table + ntab;
    m.tH = table + 2 * ntab;
    m.ntab = ntab;
    double e1[6], e0[6], s0[6], ep0[6], so[6], epo[6], nn[6], pn, A, B, gamma, rs;
    bool fl, fa;
    int it;
    for (int c = 0; c < 6; ++c) {
      e1[c] = eps[i * 6 + c];
      e0[c] = e_old[i * 6 + c];
      s0[c] = s_old[i * 6 + c];
      ep0[c] = ep_old[i * 6 + c];
    }
    dxm::j2_point<HARD, false>(m, e1, e0, s0, p_old[i], ep0, so, pn, epo, nn, A, B, gamma, fl, it, rs, fa, 1u,
                               vote != 0, true, nullptr);
    for (int c = 0; c < 6; ++c) {
      sig[i * 6 + c] = so[c];
      epsp[i * 6 + c] = epo[c];
    }
    p[i] = pn;
    // the kernel forms the 21 unique entries (j <= i) at store time; the boundary transpose mirrors them
    for (int j = 0; j < 6; ++j)
      for (int k = j; k < 6; ++k) {
        const double v = dxm::j2_tangent_entry(k, j, A, B, gamma, nn[k], nn[j]);
        ct[i * 36 + j * 6 + k] = v;
        ct[i * 36 + k * 6 + j] = v;
      }
    flag[i] = fl;
    n_iter[i] = it;
    resid[i] = rs;
    fail[i] = fa;
  }
}

}  // namespace

// hard: dxm::Hardening (0 none, 1 linear, 2 general, 3 table) -- the instantiation launch_update would pick
extern "C" void j2_host(int64_t n, const double* eps, const double* e_old, const double* s_old, const double* p_old,
                        const double* ep_old, const double* props, int perpoint, int hard, const double* table,
                        int ntab, int vote, double* sig, double* p, double* epsp, double* ct, uint8_t* flag,
                        int32_t* n_iter, double* resid, uint8_t* fail) {
#define DXM_ARGS n, eps, e_old, s_old, p_old, ep_old, props, perpoint, table, ntab, vote, sig, p, epsp, ct, flag, n_iter, resid, fail
  switch (hard) {
    case dxm::HARD_NONE: return run_j2<dxm::HARD_NONE>(DXM_ARGS);
    case dxm::HARD_LINEAR: return run_j2<dxm::HARD_LINEAR>(DXM_ARGS);
    case dxm::HARD_TABLE: return run_j2<dxm::HARD_TABLE>(DXM_ARGS);
    default: return run_j2<dxm::HARD_GENERAL>(DXM_ARGS);
  }
#undef DXM_ARGS
}

// FeFp: the point routine does its own loads and stores on SoA buffers, exactly as in the kernel.  All arrays here are
// SoA [rows][n] (the resident layout); props as above.
extern "C" void fefp_host(int64_t n, const double* F, const double* F_old, const double* p_old, const double* be_old,
                          const double* props, int perpoint, int vote, double* P, double* p, double* be, double* ct,
                          uint8_t* flag, int32_t* n_iter, double* resid, uint8_t* fail, uint64_t* n_plastic,
                          uint64_t* n_fail) {
  dxm::FeFpArgs a{};
  a.F = F;
  a.P = P;
  a.p = p;
  a.be = be;
  a.ct = ct;
  a.F_old = F_old;
  a.p_old = p_old;
  a.be_old = be_old;
  a.ld = n;
  a.start = 0;
  a.count = n;
  a.perpoint = perpoint != 0;
  if (perpoint) {
    for (int i = 0; i < 6; ++i) a.pp[i] = props + (int64_t)i * n;
  } else {  // same expressions as launch_update (dxm_api.cu)
    const double E = props[0], nu = props[1];
    a.E = E;
    a.mu = E / 2 / (1 + nu);
    a.kappa = E / (3 * (1 - 2 * nu));
    a.sig0 = props[2];
    a.H = props[3];
    const double d = props[4] - props[2];
    a.dsu = std::isfinite(d) ? d : 0.0;
    a.b = props[5];
  }
  a.vote = vote;
  a.d_flag = flag;
  a.d_iter = n_iter;
  a.d_resid = resid;
  a.d_fail = fail;
  dxm::PointStats acc;
  for (int64_t i = 0; i < n; ++i) {
    if (perpoint)
      dxm::fefp_point<true, true, false>(a, i, true, 1u, nullptr, acc);
    else
      dxm::fefp_point<false, true, false>(a, i, true, 1u, nullptr, acc);
  }
  *n_plastic = acc.n_plastic;
  *n_fail = acc.n_fail;
}
