/*
 * dxm_oracle.c -- plain-C restatement of the batched constitutive update.  TEST INFRASTRUCTURE ONLY
 * (see oracle/__init__.py): a second CPU implementation, independent of numpy, of the same canonical
 * arithmetic as oracle/small_strain.py and oracle/fefp.py; used by tests (bit-exact cross-check of the
 * numpy oracle, fast checker at large sizes) and by bench.py's CPU baseline.  Never linked into the product.
 *
 * Restates: Material.integrate (dolfinx_materials/generic.py:176-189, jaxmat.py:208-234);
 * elasticity python_materials/elasticity.py:12-24; J2 closed form
 * tests/mfront/IsotropicLinearHardeningPlasticity.mfront:49-77; Voce law tests/test_FeFp_jax.py:14-15;
 * FeFp per SURVEY.md A.4 (parity unpinned: jaxmat is un-vendored).
 *
 * Build (oracle/Makefile): gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC   (single-threaded C;
 * oracle/cport.py runs row blocks on Python threads -- ctypes releases the GIL -- since this image has no libgomp)
 *   -ffp-contract=off is essential: every operation is individually rounded, in the order written.
 * Arrays are C-contiguous (n, dim) float64 exactly as the reference's QuadratureMap hands them over.
 */
#include "dxm_canon.h"

#define NEWTON_CAP_DEFAULT 25

static double rcbrt_c(double x) {
  if (!(x > 0.0)) return NAN;
  int e;
  const double m = frexp(x, &e);
  const int q = (e >= 0) ? (e / 3) : -((-e + 2) / 3);
  const int r = e - 3 * q;
  const double xr = ldexp(m, r);
  double y = FNMA(0.15, xr, 1.2);
  for (int i = 0; i < 6; ++i) y = (y * FNMA(xr, (y * y) * y, 4.0)) * (1.0 / 3.0);
  return ldexp(y, -q);
}

typedef struct {
  double E, nu, sig0, H, sigu, b;
} props_t;

static props_t get_props(const double* const* pp, const int* per, int64_t i) {
  props_t m;
  m.E = per[0] ? pp[0][i] : pp[0][0];
  m.nu = per[1] ? pp[1][i] : pp[1][0];
  m.sig0 = per[2] ? pp[2][i] : pp[2][0];
  m.H = per[3] ? pp[3][i] : pp[3][0];
  m.sigu = per[4] ? pp[4][i] : pp[4][0];
  m.b = per[5] ? pp[5][i] : pp[5][0];
  return m;
}

/* small strain: elastic (sig0 = inf) / J2 linear / Voce / mixed -- order of operations == oracle/small_strain.py */
void dxo_small_strain(int64_t n, const double* eps, const double* e_old, const double* s_old, const double* p_old_a,
                      const double* ep_old, const double* const* pp, const int* per, int newton_cap, double rtol,
                      double* sig_o, double* p_o, double* epsp_o, double* ct_o, uint8_t* flag_o, int32_t* iter_o,
                      double* resid_o, uint8_t* fail_o) {
  for (int64_t pt = 0; pt < n; ++pt) {
    const props_t m = get_props(pp, per, pt);
    const double lam = m.E * m.nu / (1 + m.nu) / (1 - 2 * m.nu);
    const double mu = m.E / 2 / (1 + m.nu);
    const double twomu = 2.0 * mu, threemu = 3.0 * mu;
    double dsu = m.sigu - m.sig0;
    if (!isfinite(dsu)) dsu = 0.0;
    const double bdsu = m.b * dsu;
    const double p_old = p_old_a[pt];
    double de[6], st[6], s[6];
    for (int i = 0; i < 6; ++i) de[i] = eps[pt * 6 + i] - e_old[pt * 6 + i];
    const double tr = (de[0] + de[1]) + de[2];
    const double ltr = lam * tr;
    for (int i = 0; i < 3; ++i) st[i] = s_old[pt * 6 + i] + FMA(twomu, de[i], ltr);
    for (int i = 3; i < 6; ++i) st[i] = FMA(twomu, de[i], s_old[pt * 6 + i]);
    const double pm = ((st[0] + st[1]) + st[2]) / 3.0;
    for (int i = 0; i < 3; ++i) s[i] = st[i] - pm;
    for (int i = 3; i < 6; ++i) s[i] = st[i];
    double ss = s[0] * s[0];
    for (int i = 1; i < 6; ++i) ss = FMA(s[i], s[i], ss);
    const double seq = sqrt(1.5 * ss);
    double ecur = exp_c(-(m.b * p_old));
    const double sy0 = FMA(dsu, 1.0 - ecur, FMA(m.H, p_old, m.sig0));
    const double f = seq - sy0;
    const int flag = f > 0.0;
    double dp = 0.0, resid = 0.0;
    int n_iter = 0, fail = 0;
    if (flag) {
      if (bdsu == 0.0) {
        dp = f / (threemu + m.H);
      } else {
        const double tol = rtol * seq;
        for (int it = 0;; ++it) {
          const double p = p_old + dp;
          const double sy = FMA(dsu, 1.0 - ecur, FMA(m.H, p, m.sig0));
          const double r = FNMA(threemu, dp, seq) - sy;
          if (fabs(r) <= tol) { resid = fabs(r); break; }
          if (it == newton_cap) { resid = fabs(r); fail = 1; break; }
          const double dsy = FMA(bdsu, ecur, m.H);
          dp = dp + r / (threemu + dsy);
          ecur = exp_c(-(m.b * (p_old + dp)));
          ++n_iter;
        }
      }
    }
    const double Hp = FMA(bdsu, ecur, m.H);
    double nrm[6], q = 0.0, gamma = 0.0;
    if (flag) {
      for (int i = 0; i < 6; ++i) nrm[i] = (1.5 * s[i]) / seq;
      q = dp / seq;
    } else {
      for (int i = 0; i < 6; ++i) nrm[i] = 0.0;
      dp = 0.0;
    }
    double epsp[6];
    for (int i = 0; i < 6; ++i) {
      const double depsp = dp * nrm[i];
      sig_o[pt * 6 + i] = FNMA(twomu, depsp, st[i]);
      epsp[i] = ep_old[pt * 6 + i] + depsp;
      epsp_o[pt * 6 + i] = epsp[i];
    }
    const double p_new = p_old + dp;
    p_o[pt] = p_new;
    const double fourmu2 = (4.0 * mu) * mu;
    const double beta = fourmu2 * q;
    if (flag) {
      const double cste = 1.0 / (threemu + Hp);
      gamma = fourmu2 * (cste - q);
    }
    const double A = FMA(0.5, beta, lam), B = FNMA(1.5, beta, twomu), AB = A + B;
    for (int j = 0; j < 6; ++j)
      for (int i = 0; i < 6; ++i) {
        const double base = (i == j) ? ((i < 3) ? AB : B) : ((i < 3 && j < 3) ? A : 0.0);
        ct_o[pt * 36 + j * 6 + i] = FNMA(gamma, nrm[i] * nrm[j], base);
      }
    double chk = (seq + fabs(pm)) + p_new;
    for (int i = 0; i < 6; ++i) chk = chk + fabs(epsp[i]);
    if (!isfinite(chk)) fail = 1;
    flag_o[pt] = (uint8_t)flag;
    iter_o[pt] = n_iter;
    resid_o[pt] = resid;
    fail_o[pt] = (uint8_t)fail;
  }
}

/* ---- finite strain FeFp -- order of operations == oracle/fefp.py ------------------------------------ */
static const int IDX9[3][3] = {{0, 3, 5}, {4, 1, 7}, {6, 8, 2}};

static double det3(double A[3][3]) {
  const double m0 = FMS(A[1][1], A[2][2], A[1][2] * A[2][1]);
  const double m1 = FMS(A[1][0], A[2][2], A[1][2] * A[2][0]);
  const double m2 = FMS(A[1][0], A[2][1], A[1][1] * A[2][0]);
  return FMA(A[0][2], m2, FNMA(A[0][1], m1, A[0][0] * m0));
}

static double inv3(double A[3][3], double Ai[3][3]) {
  double c[3][3];
  c[0][0] = FMS(A[1][1], A[2][2], A[1][2] * A[2][1]);
  c[0][1] = FMS(A[0][2], A[2][1], A[0][1] * A[2][2]);
  c[0][2] = FMS(A[0][1], A[1][2], A[0][2] * A[1][1]);
  c[1][0] = FMS(A[1][2], A[2][0], A[1][0] * A[2][2]);
  c[1][1] = FMS(A[0][0], A[2][2], A[0][2] * A[2][0]);
  c[1][2] = FMS(A[0][2], A[1][0], A[0][0] * A[1][2]);
  c[2][0] = FMS(A[1][0], A[2][1], A[1][1] * A[2][0]);
  c[2][1] = FMS(A[0][1], A[2][0], A[0][0] * A[2][1]);
  c[2][2] = FMS(A[0][0], A[1][1], A[0][1] * A[1][0]);
  const double det = FMA(A[0][2], c[2][0], FMA(A[0][1], c[1][0], A[0][0] * c[0][0]));
  const double rdet = 1.0 / det;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Ai[i][j] = c[i][j] * rdet;
  return det;
}

void dxo_fefp(int64_t n, const double* F, const double* F_old, const double* p_old_a, const double* be_old,
              const double* const* pp, const int* per, int newton_cap, double rtol, double* P_o, double* p_o,
              double* be_o, double* ct_o, uint8_t* flag_o, int32_t* iter_o, double* resid_o, uint8_t* fail_o) {
  const double RSQRT2 = 0.70710678118654752440, SQRT2 = 1.41421356237309504880, THIRD = 1.0 / 3.0;
  for (int64_t pt = 0; pt < n; ++pt) {
    const props_t m = get_props(pp, per, pt);
    const double mu = m.E / 2 / (1 + m.nu);
    const double kappa = m.E / (3 * (1 - 2 * m.nu));
    const double threemu = 3.0 * mu;
    double dsu = m.sigu - m.sig0;
    if (!isfinite(dsu)) dsu = 0.0;
    const double bdsu = m.b * dsu;
    const double p_old = p_old_a[pt];
    double A[3][3], Ao[3][3], Bo[3][3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        A[i][j] = F[pt * 9 + IDX9[i][j]];
        Ao[i][j] = F_old[pt * 9 + IDX9[i][j]];
      }
    const double* beo = be_old + pt * 6;
    Bo[0][0] = beo[0]; Bo[1][1] = beo[1]; Bo[2][2] = beo[2];
    Bo[0][1] = Bo[1][0] = beo[3] * RSQRT2;
    Bo[0][2] = Bo[2][0] = beo[4] * RSQRT2;
    Bo[1][2] = Bo[2][1] = beo[5] * RSQRT2;
    const double det_bo = det3(Bo); /* a singular / inverted elastic state is a failed point */
    double Aoi[3][3], f[3][3], M[3][3], B[3][3], D[3][3];
    inv3(Ao, Aoi);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) f[i][j] = dot3(A[i][0], Aoi[0][j], A[i][1], Aoi[1][j], A[i][2], Aoi[2][j]);
    const double Jf = det3(f);
    const double rc = rcbrt_c(Jf);
    const double s23 = rc * rc;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) M[i][j] = dot3(f[i][0], Bo[0][j], f[i][1], Bo[1][j], f[i][2], Bo[2][j]);
    for (int i = 0; i < 3; ++i)
      for (int j = i; j < 3; ++j) {
        B[i][j] = s23 * dot3(M[i][0], f[j][0], M[i][1], f[j][1], M[i][2], f[j][2]);
        B[j][i] = B[i][j];
      }
    const double t0 = ((B[0][0] + B[1][1]) + B[2][2]) * THIRD;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) D[i][j] = (i == j) ? (B[i][j] - t0) : B[i][j];
    const double dd = FMA(2.0, dot3(D[0][1], D[0][1], D[0][2], D[0][2], D[1][2], D[1][2]),
                          dot3(D[0][0], D[0][0], D[1][1], D[1][1], D[2][2], D[2][2]));
    const double d3 = det3(D);
    const double seq = mu * sqrt(1.5 * dd);
    const double rseq = 1.0 / seq;
    double ecur = exp_c(-(m.b * p_old));
    const double sy0 = FMA(dsu, 1.0 - ecur, FMA(m.H, p_old, m.sig0));
    const int flag = (seq - sy0) > 0.0;
    const double c = threemu * rseq;
    double dp = 0.0, t = t0, resid = 0.0;
    int n_iter = 0, fail = 0;
    if (flag) {
      const double tol1 = rtol * seq;
      for (int it = 0;; ++it) {
        const double ct = c * t;
        const double tmt = threemu * t;
        const double alpha = FNMA(ct, dp, 1.0);
        const double p = p_old + dp;
        const double sy = FMA(dsu, 1.0 - ecur, FMA(m.H, p, m.sig0));
        const double r1 = FNMA(tmt, dp, seq) - sy;
        const double a2 = alpha * alpha;
        const double ha2 = 0.5 * a2;
        const double tt = t * t;
        const double r2 = FNMA(ha2, dd * t, tt * t) + FMS(a2 * alpha, d3, 1.0);
        if (fabs(r1) <= tol1 && fabs(r2) <= rtol) { resid = fabs(r1); break; }
        if (it == newton_cap) { resid = fabs(r1); fail = 1; break; }
        const double dsy = FMA(bdsu, ecur, m.H);
        const double g = FNMA(alpha * dd, t, 3.0 * (a2 * d3));
        const double J11 = -tmt - dsy, J12 = -(threemu * dp);
        const double cdp = c * dp;
        const double J21 = -(g * ct);
        const double J22 = FNMA(g, cdp, FNMA(ha2, dd, 3.0 * tt));
        const double rdet = 1.0 / FMS(J11, J22, J12 * J21);
        const double dp_new = FMA(FMS(J12, r2, r1 * J22), rdet, dp);
        const double t_new = FMA(FMS(J21, r1, J11 * r2), rdet, t);
        dp = dp_new;
        t = t_new;
        ecur = exp_c(-(m.b * (p_old + dp)));
        ++n_iter;
      }
    }
    const double alpha = flag ? FNMA(c * t, dp, 1.0) : 1.0;
    const double p_new = p_old + dp;
    double be[6];
    for (int i = 0; i < 3; ++i) be[i] = flag ? FMA(alpha, D[i][i], t) : B[i][i];
    be[3] = (alpha * D[0][1]) * SQRT2;
    be[4] = (alpha * D[0][2]) * SQRT2;
    be[5] = (alpha * D[1][2]) * SQRT2;
    double Ai[3][3], DA[3][3], P[3][3];
    const double Jd = inv3(A, Ai);
    const double muA = mu * alpha;
    const double pvol = (0.5 * kappa) * FMS(Jd, Jd, 1.0);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) DA[i][j] = dot3(D[i][0], Ai[j][0], D[i][1], Ai[j][1], D[i][2], Ai[j][2]);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) P[i][j] = FMA(muA, DA[i][j], pvol * Ai[j][i]);
    double al1 = 0.0, al2 = 0.0;
    if (flag) {
      const double sq1 = (1.5 * (mu * mu)) * rseq;
      const double a2 = alpha * alpha;
      const double dsy = FMA(bdsu, ecur, m.H);
      const double g = FNMA(alpha * dd, t, 3.0 * (a2 * d3));
      const double ct = c * t, cdp = c * dp;
      const double J11 = -(threemu * t) - dsy, J12 = -(threemu * dp);
      const double J21 = -(g * ct);
      const double J22 = FNMA(g, cdp, FNMA(0.5 * a2, dd, 3.0 * (t * t)));
      const double rdet = 1.0 / FMS(J11, J22, J12 * J21);
      const double oma = (1.0 - alpha) * rseq;
      const double b21 = FMS(g * oma, sq1, a2 * t);
      const double b22 = a2 * alpha;
      const double p1 = -(FMS(sq1, J22, J12 * b21) * rdet);
      const double t1 = -(FMS(J11, b21, J21 * sq1) * rdet);
      const double p2 = (J12 * b22) * rdet;
      const double t2 = -((J11 * b22) * rdet);
      al1 = FNMA(cdp, t1, FNMA(ct, p1, oma * sq1));
      al2 = FNMA(cdp, t2, -(ct * p2));
    }
    const double kJ2 = kappa * (Jd * Jd);
    const double c23dd = (2.0 / 3.0) * dd, twod3 = 2.0 * d3, c23muA = (2.0 / 3.0) * muA;
    const double hs = FMS(muA, t0, pvol);
    for (int l = 0; l < 3; ++l) {
      double w[3], v[3], u[3], z[3], my[3], hw[3];
      for (int i = 0; i < 3; ++i) w[i] = Ai[l][i];
      for (int i = 0; i < 3; ++i) v[i] = FMA(t0, w[i], dot3(D[i][0], w[0], D[i][1], w[1], D[i][2], w[2]));
      for (int i = 0; i < 3; ++i) u[i] = dot3(D[i][0], v[0], D[i][1], v[1], D[i][2], v[2]);
      for (int i = 0; i < 3; ++i) z[i] = dot3(D[i][0], u[0], D[i][1], u[1], D[i][2], u[2]);
      for (int j = 0; j < 3; ++j) my[j] = muA * dot3(Ai[j][0], v[0], Ai[j][1], v[1], Ai[j][2], v[2]);
      for (int i = 0; i < 3; ++i) hw[i] = hs * w[i];
      for (int k = 0; k < 3; ++k) {
        double cd0 = 0.0;
        if (flag) {
          const double a1 = FNMA(c23dd, w[k], 2.0 * u[k]);
          const double a2p = FNMA(twod3, w[k], FNMA(c23dd, v[k], 2.0 * z[k]));
          cd0 = mu * FMA(al2, a2p, al1 * a1);
        }
        const double cD = FNMA(c23muA, w[k], cd0);
        const double cI = FNMA(c23muA, v[k], kJ2 * w[k]);
        const int col = IDX9[k][l];
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j) {
            double val = FMA(hw[i], Ai[j][k], FMA(cI, Ai[j][i], cD * DA[i][j]));
            if (i == k) val = val + my[j];
            ct_o[pt * 81 + IDX9[i][j] * 9 + col] = val;
          }
      }
    }
    double chk = (seq + fabs(Jd)) + p_new;
    for (int i = 0; i < 6; ++i) chk = chk + fabs(be[i]);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) chk = chk + fabs(P[i][j]);
    if (!isfinite(chk) || !(det_bo > 0.0)) fail = 1;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) P_o[pt * 9 + IDX9[i][j]] = P[i][j];
    p_o[pt] = p_new;
    for (int i = 0; i < 6; ++i) be_o[pt * 6 + i] = be[i];
    flag_o[pt] = (uint8_t)flag;
    iter_o[pt] = n_iter;
    resid_o[pt] = resid;
    fail_o[pt] = (uint8_t)fail;
  }
}
