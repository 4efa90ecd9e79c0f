// Small-strain plasticity with the Hosford criterion and isotropic hardening -- linear hardening is the behaviour of the
// matrix phase of the reference's multi-material demo (demos/multimaterials/IsotropicPlasticHosfordFlowLinear.mfront:1-27,
// used at demos/multimaterials/multimaterials.py:245-254 through MFront/MGIS); with the general law of the J2 kernels
// (sig0 + H p + (sigu - sig0)(1 - exp(-b p))) it is jaxmat's GeneralIsotropicHardening(elastic, yield_stress, ...)
// (demos/jax/elastoplasticity/_plane_stress_elastoplasticity.py:17,45) -- behind the same handle / state layout
// (strain, stress, p, epsp) as the J2 behaviours.
//
//   sigma_eq = (1/2 (|s1-s2|^a + |s2-s3|^a + |s3-s1|^a))^(1/a),  a even integer (2 = von Mises, 10 in the demo)
//   backward Euler, associated flow:  sigma = sigma_tr - 2 mu dp n(sigma),  sigma_eq(sigma) = sigma_Y(p_old + dp)
//
// One Gauss point per thread.  Isotropy keeps the principal axes of the trial stress, so per point:
//   * cheap rejection: sigma_eq <= max|s_i - s_j| <= 2/sqrt(3) seq_Mises  -> clearly elastic points skip the rest;
//   * non-iterative eigen-decomposition of the trial deviator in registers (hos_eig3: isolated root of the
//     characteristic cubic, cross product, one Jacobi rotation in the normal plane; + - * / sqrt only); the equivalent
//     stress costs one division per evaluation (the a-th root is a division-free fixed-count iteration on q^(-1/a));
//   * 4-unknown Newton (3 principal deviatoric stresses + dp) from the radially scaled trial state, with a
//     simple-decrease backtracking line search (plain Newton overshoots at the rounded corners of the surface);
//     3x3 solves by the symmetric adjugate;
//   * consistent tangent from its spectral form: normal block 2 mu A^-1 + lam 1 1^T - w z z^T and three shear moduli
//     2 mu / (1 + 2 mu dp theta_ij), theta_ij = (n_i - n_j)/(s_i - s_j) as an exact divided difference (even a), rotated
//     back with the eigenvectors and stored packed (21 unique entries) like the J2 tangent.
// All powers are product chains; operation order == oracle/c/dxm_oracle_hosford.c (bit-identical with -fmad=false).
// The point routine is __host__ __device__ so that a CPU test can run the very same code against the oracle
// (tests/hosford_host_check.cu) -- the product only ever calls it from the kernels below.
//
// Launch structure:
//   dxm_hosford_kernel        fused: every thread runs the full routine on its own point -- the default at every plastic
//                             fraction since the eigen-decomposition became non-iterative (profiles/r02t_hosford_ab.json);
//   dxm_hosford_tiled_kernel  each CTA streams a 1024-point tile, finishes the clearly elastic points and packs the
//                             candidates into full warps through a shared-memory queue -- won below ~30 % plastic while
//                             a warp with ONE candidate lane paid a twice as expensive local solve
//                             (profiles/r01f_hosford_v1_ncu_*); now an opt-in (DXM_HOS_SPLIT=1).
#pragma once
#include "dxm_canon.cuh"
#include "dxm_small_strain.cuh"

namespace dxm {

constexpr double kHosRSqrt2 = 0.7071067811865476;
constexpr double kHosSqrt2 = 1.4142135623730951;
constexpr int kHosfordLsMax = 10;

// exp_c written for host and device: exp_hd of dxm_canon.cuh
DXM_HD double hos_exp(double x) { return exp_hd(x); }

// isotropic hardening sigma_Y(p) = sig0 + H p + dsu (1 - exp(-b p)) and its slope at p = p_old + dp (the law of the J2
// behaviours: linear for dsu = 0, Voce for H = 0)
struct HosHard {
  double sig0, H, dsu, b, bdsu, p_old;
};

DXM_HD void hos_hard(const HosHard& hd, double dp, double& sy, double& dsy) {
  const double p = hd.p_old + dp;
  const double e = (hd.bdsu != 0.0) ? hos_exp(-(hd.b * p)) : 1.0;
  sy = fma_c(hd.dsu, 1.0 - e, fma_c(hd.H, p, hd.sig0));
  dsy = fma_c(hd.bdsu, e, hd.H);
}

// (x*x)^k, 1 <= k <= 32, by binary powering from the top bit of k: y = x^2; per lower bit: y = y*y, then y = y * x^2 if
// the bit is set -- k = 5 (a = 10): x^2, x^4, x^8, x^10: 4 products, depth 4 (the round-1 chain multiplied k-1 times by
// x^2: depth k).  With a compile-time exponent the loop folds into straight DMULs.
DXM_HD double hos_ipow2(double x, int k) {
  const double x2 = x * x;
  double y = x2;
  int top = 5;
  while (top > 0 && !((k >> top) & 1)) --top;
#pragma unroll
  for (int bit = 4; bit >= 0; --bit) {
    if (bit < top) {
      y = y * y;
      if ((k >> bit) & 1) y = y * x2;
    }
  }
  return y;
}

// q^(-1/a), q in (0.5, 1], division free: second-order Taylor start in x = 1 - q (relative error <= 5 % for a = 2,
// 0.8 % for a = 10), then a FIXED two steps of the third-order correction: with r = 1 - q w^a the exact root is
// w (1 - r)^(-1/a) = w (1 + s r (1 + (s+1)/2 r (1 + (s+2)/3 r (...)))), s = 1/a; truncated after r^3 the error goes
// e -> O(e^4): below 2e-18 after two steps for every even a in [2, 64] (scanned in tests/test_oracle_hosford.py).
// History: round 1 started from w = 1 and ran plain Newton steps until rounding stopped the monotone sequence (5-7
// data-dependent trips on the critical path of every evaluation of the criterion); then a fixed four Newton steps
// (chain depth 3 + 4 x 7); now 3 + 2 x 9.
constexpr int kHosRootSteps = 2;
DXM_HD double hos_arootinv(double q, int a, double inv_a) {
  const double x = 1.0 - q;
  const double k2 = 0.5 * (1.0 + inv_a);
  const double k3 = (2.0 + inv_a) / 3.0;
  double w = fma_c(x * inv_a, fma_c(x, k2, 1.0), 1.0);
#pragma unroll
  for (int it = 0; it < kHosRootSteps; ++it) {
    const double r = fnma_c(q, hos_ipow2(w, a / 2), 1.0);
    w = w * fma_c(r * inv_a, fma_c(r * k2, fma_c(r, k3, 1.0), 1.0), 1.0);
  }
  return w;
}

struct HosEval {
  double phi, iphi, n[3], h[3], u[3];
};

DXM_HD void hos_eval(const double (&l)[3], int a, double inv_a, HosEval& e) {
  const double d0 = l[0] - l[1], d1 = l[1] - l[2], d2 = l[2] - l[0];
  const double m = fmax(fmax(fabs(d0), fabs(d1)), fabs(d2));
  const double im = 1.0 / m;
  const double r0 = d0 * im, r1 = d1 * im, r2 = d2 * im;
  const double q = 0.5 * ((hos_ipow2(r0, a / 2) + hos_ipow2(r1, a / 2)) + hos_ipow2(r2, a / 2));
  const double w = hos_arootinv(q, a, inv_a);
  // phi = m q^(1/a) = m q w^(a-1): a short product chain instead of the division m / w (relative error <= ~a ulp,
  // three orders of magnitude below the Newton tolerance)
  e.phi = m * (q * ((a > 2) ? hos_ipow2(w, (a - 2) / 2) * w : w));
  e.iphi = w * im;
  e.u[0] = r0 * w;
  e.u[1] = r1 * w;
  e.u[2] = r2 * w;
#pragma unroll
  for (int k = 0; k < 3; ++k) e.h[k] = (a > 2) ? hos_ipow2(e.u[k], (a - 2) / 2) : 1.0;
  const double g0 = e.h[0] * e.u[0], g1 = e.h[1] * e.u[1], g2 = e.h[2] * e.u[2];
  e.n[0] = 0.5 * (g0 - g2);
  e.n[1] = 0.5 * (g1 - g0);
  e.n[2] = 0.5 * (g2 - g1);
}

// sum_{k=0}^{a-2} x^k y^(a-2-k)
DXM_HD double hos_divdiff(double x, double y, int a) {
  double t = 1.0, xp = 1.0;
#pragma unroll
  for (int j = 1; j <= a - 2; ++j) {
    xp = xp * x;
    t = fma_c(y, t, xp);
  }
  return t;
}

// ---- symmetric 3x3 eigen-decomposition of a deviator, non-iterative -------------------------------------------------
// (replaces the cyclic Jacobi sweeps of round 1 / early round 2: 4-5 sweeps x 3 rotations per warp, each rotation two
// square roots and two divisions -- 19 % of the fused kernel's time at 80 % plastic points, profiles/r02r_*.)  With
// B = A / p, p = sqrt(tr(A^2) / 6), the eigenvalues of B are 2 cos(theta + 2 pi j / 3), cos(3 theta) = det(B) / 2 =: k.
// The eigenvalue on the side of the sign of k is ISOLATED (at least 0.866 * 2 from the other two, whatever the spectrum):
//   1. y = cos(theta) in [sqrt(3)/2, 1] solves 4 y^3 - 3 y = |k|: cubic start polynomial + three division-free Newton
//      steps (the reciprocal slope R is refined alongside, R <- R (2 - g' R)); beta0 = sgn(k) 2 y;
//   2. its eigenvector n = the largest of the three cross products of rows of B - beta0 I, normalised (well conditioned
//      because beta0 is isolated);
//   3. an orthonormal basis (U, W) of the plane normal to n without a square root (Duff et al., "Building an orthonormal
//      basis, revisited", JCGT 2017) and ONE exact Jacobi rotation of the 2x2 restriction of B to that plane; repeated or
//      nearly repeated eigenvalues there are harmless (any rotation of an eigenplane is a valid basis).
// Eigenvalues are Rayleigh quotients of the computed vectors.  Straight-line code, no data-dependent trip count; residual
// |A V - V L| / |A| and |V^T V - I| <= ~1.5e-15 over random, axisymmetric, nearly degenerate and pure-shear spectra
// (tests/test_oracle_hosford.py on the bit-identical C twin).
constexpr double kEigCY0 = 0.8660615980506479, kEigCY1 = 0.16540585304875938, kEigCY2 = -0.04088323804684024,
                 kEigCY3 = 0.009444179663245256;
constexpr double kEigCR0 = 0.1660512323983123, kEigCR1 = -0.08345496691468178, kEigCR2 = 0.028968785247117986;
constexpr double kEigSixth = 0.16666666666666666;

DXM_HD void hos_cross3(const double (&u)[3], const double (&v)[3], double (&c)[3], double& d) {
  c[0] = fms_c(u[1], v[2], u[2] * v[1]);
  c[1] = fms_c(u[2], v[0], u[0] * v[2]);
  c[2] = fms_c(u[0], v[1], u[1] * v[0]);
  d = fma_c(c[2], c[2], fma_c(c[1], c[1], c[0] * c[0]));
}

DXM_HD double hos_dot3(const double (&u)[3], const double (&v)[3]) { return fma_c(u[2], v[2], fma_c(u[1], v[1], u[0] * v[0])); }

// s: Mandel deviator; l: eigenvalues; V: eigenvectors in the columns
DXM_HD void hos_eig3(const double (&s)[6], double (&l)[3], double (&V)[3][3]) {
  const double a00 = s[0], a11 = s[1], a22 = s[2], a01 = s[3] * kHosRSqrt2, a02 = s[4] * kHosRSqrt2, a12 = s[5] * kHosRSqrt2;
  const double dg = fma_c(a22, a22, fma_c(a11, a11, a00 * a00));
  const double od = fma_c(a12, a12, fma_c(a02, a02, a01 * a01));
  const double p2 = fma_c(2.0, od, dg) * kEigSixth;
  if (p2 == 0.0) {  // zero deviator (never a candidate point)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      l[i] = 0.0;
#pragma unroll
      for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    }
    return;
  }
  const double p = sqrt(p2), ip = 1.0 / p;
  const double b00 = a00 * ip, b11 = a11 * ip, b22 = a22 * ip, b01 = a01 * ip, b02 = a02 * ip, b12 = a12 * ip;
  const double m0 = fms_c(b11, b22, b12 * b12), m1 = fms_c(b01, b22, b12 * b02), m2 = fms_c(b01, b12, b11 * b02);
  const double hdet = 0.5 * fma_c(b02, m2, fms_c(b00, m0, b01 * m1));
  const double sgn = (hdet >= 0.0) ? 1.0 : -1.0;
  const double k = fmin(fabs(hdet), 1.0);
  double y = fma_c(fma_c(fma_c(kEigCY3, k, kEigCY2), k, kEigCY1), k, kEigCY0);
  double R = fma_c(fma_c(kEigCR2, k, kEigCR1), k, kEigCR0);
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    const double y2 = y * y;
    const double g = fms_c(fms_c(4.0, y2, 3.0), y, k);
    const double gp = fms_c(12.0, y2, 3.0);
    R = R * fnma_c(gp, R, 2.0);
    y = fnma_c(g, R, y);
  }
  const double beta = sgn * (2.0 * y);
  const double r0[3] = {b00 - beta, b01, b02}, r1[3] = {b01, b11 - beta, b12}, r2[3] = {b02, b12, b22 - beta};
  double c[3], d, c2[3], d2;
  hos_cross3(r0, r1, c, d);
  hos_cross3(r0, r2, c2, d2);
  if (d2 > d) {
    c[0] = c2[0];
    c[1] = c2[1];
    c[2] = c2[2];
    d = d2;
  }
  hos_cross3(r1, r2, c2, d2);
  if (d2 > d) {
    c[0] = c2[0];
    c[1] = c2[1];
    c[2] = c2[2];
    d = d2;
  }
  const double inv = 1.0 / sqrt(d);
  const double n[3] = {c[0] * inv, c[1] * inv, c[2] * inv};
  const double sg = (n[2] >= 0.0) ? 1.0 : -1.0;
  const double a = -1.0 / (sg + n[2]);
  const double bq = (n[0] * n[1]) * a;
  const double U[3] = {fma_c(sg * (n[0] * n[0]), a, 1.0), sg * bq, -(sg * n[0])};
  const double W[3] = {bq, fma_c(n[1] * n[1], a, sg), -n[1]};
  double Bn[3], BU[3], BW[3];
#define DXM_EIG_MV(v, o)                                        \
  o[0] = fma_c(b02, v[2], fma_c(b01, v[1], b00 * v[0]));        \
  o[1] = fma_c(b12, v[2], fma_c(b11, v[1], b01 * v[0]));        \
  o[2] = fma_c(b22, v[2], fma_c(b12, v[1], b02 * v[0]));
  DXM_EIG_MV(n, Bn)
  DXM_EIG_MV(U, BU)
  DXM_EIG_MV(W, BW)
#undef DXM_EIG_MV
  const double l0 = hos_dot3(n, Bn), m00 = hos_dot3(U, BU), m01 = hos_dot3(U, BW), m11 = hos_dot3(W, BW);
  // the Jacobi rotation of [[m00, m01], [m01, m11]]: t = tangent of the smaller angle, one division
  const double delta = (m11 - m00) * 0.5;
  const double den = fabs(delta) + sqrt(fma_c(delta, delta, m01 * m01));
  double t = (den > 0.0) ? m01 / den : 0.0;
  if (delta < 0.0) t = -t;
  const double cs = 1.0 / sqrt(fma_c(t, t, 1.0)), sn = t * cs;
  l[0] = l0 * p;
  l[1] = fnma_c(t, m01, m00) * p;
  l[2] = fma_c(t, m01, m11) * p;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    V[i][0] = n[i];
    V[i][1] = fms_c(cs, U[i], sn * W[i]);
    V[i][2] = fma_c(sn, U[i], cs * W[i]);
  }
}

struct HosRes {
  HosEval e;
  double rs[3], r4, m2, dsy;
};

// residuals of the 4 unknowns (x, dp) from the criterion data o.e already evaluated at x
DXM_HD void hos_residual_finish(const double (&x)[3], double dp, const double (&l)[3], double twomu, const HosHard& hd,
                                HosRes& o) {
  const double c = twomu * dp;
#pragma unroll
  for (int k = 0; k < 3; ++k) o.rs[k] = fma_c(c, o.e.n[k], x[k] - l[k]);
  double sy;
  hos_hard(hd, dp, sy, o.dsy);
  o.r4 = o.e.phi - sy;
  o.m2 = fma_c(o.r4, o.r4, fma_c(o.rs[2], o.rs[2], fma_c(o.rs[1], o.rs[1], o.rs[0] * o.rs[0])));
}

// A = I + c k1 (M/2 - n n^T): adjugate (6 unique cofactors) and 1/det
DXM_HD void hos_system(const HosEval& r, double c, double k1, double (&Cf)[6], double& det) {
  const double ck = c * k1;
  const double A00 = fma_c(ck, fnma_c(r.n[0], r.n[0], 0.5 * (r.h[0] + r.h[2])), 1.0);
  const double A11 = fma_c(ck, fnma_c(r.n[1], r.n[1], 0.5 * (r.h[0] + r.h[1])), 1.0);
  const double A22 = fma_c(ck, fnma_c(r.n[2], r.n[2], 0.5 * (r.h[1] + r.h[2])), 1.0);
  const double A01 = ck * fnma_c(r.n[0], r.n[1], -0.5 * r.h[0]);
  const double A02 = ck * fnma_c(r.n[0], r.n[2], -0.5 * r.h[2]);
  const double A12 = ck * fnma_c(r.n[1], r.n[2], -0.5 * r.h[1]);
  Cf[0] = fms_c(A11, A22, A12 * A12);
  Cf[1] = fms_c(A02, A12, A01 * A22);
  Cf[2] = fms_c(A01, A12, A02 * A11);
  Cf[3] = fms_c(A00, A22, A02 * A02);
  Cf[4] = fms_c(A01, A02, A00 * A12);
  Cf[5] = fms_c(A00, A11, A01 * A01);
  det = fma_c(A02, Cf[2], fma_c(A01, Cf[1], A00 * Cf[0]));
}

// adj(A) v (the caller scales by 1/det where it needs A^-1 v)
DXM_HD void hos_apply(const double (&Cf)[6], const double (&v)[3], double (&o)[3]) {
  o[0] = fma_c(Cf[2], v[2], fma_c(Cf[1], v[1], Cf[0] * v[0]));
  o[1] = fma_c(Cf[4], v[2], fma_c(Cf[3], v[1], Cf[1] * v[0]));
  o[2] = fma_c(Cf[5], v[2], fma_c(Cf[4], v[1], Cf[2] * v[0]));
}

// unit Mandel vector of sym(e_I e_J) (I != J) or of e_I e_I, from the eigenvector matrix
template <int I, int J>
DXM_HD void hos_mandel_pair(const double (&V)[3][3], double (&m)[6]) {
  if (I == J) {
    m[0] = V[0][I] * V[0][I];
    m[1] = V[1][I] * V[1][I];
    m[2] = V[2][I] * V[2][I];
    m[3] = kHosSqrt2 * (V[0][I] * V[1][I]);
    m[4] = kHosSqrt2 * (V[0][I] * V[2][I]);
    m[5] = kHosSqrt2 * (V[1][I] * V[2][I]);
  } else {
    m[0] = kHosSqrt2 * (V[0][I] * V[0][J]);
    m[1] = kHosSqrt2 * (V[1][I] * V[1][J]);
    m[2] = kHosSqrt2 * (V[2][I] * V[2][J]);
    m[3] = fma_c(V[0][I], V[1][J], V[1][I] * V[0][J]);
    m[4] = fma_c(V[0][I], V[2][J], V[2][I] * V[0][J]);
    m[5] = fma_c(V[1][I], V[2][J], V[2][I] * V[1][J]);
  }
}

// ---- one Gauss point, in four stages ---------------------------------------------------------------------------------
// hos_trial -> hos_eig3 -> hos_newton -> hos_finish.  hosford_point() below runs them back to back (CPU harness,
// phase A of the tiled kernel); the local-solve kernels run the same stages but keep only what the Newton loop needs in
// registers across it: the eigenvectors wait in shared memory and the trial stress / old plastic strain are re-formed
// from the (L2-resident) inputs afterwards -- same operations on the same values, hence the same bits, with 128 instead
// of 168 registers per thread (profiles/r02c_hosford_*).
//
// AT > 0: the exponent as a compile-time constant (the power chains unroll into straight DMUL sequences; with a
// run-time exponent 35 % of the executed instructions were loop bookkeeping, profiles/r01g_hosford_v2_*); AT == 0: a_rt.
// bound: (2^(a-1)+1)^(1/a)/sqrt(3) (1 + 1e-9) >= sigma_eq / seq_Mises for every stress state (maximum at pure shear).
// VOCE == false: the saturation term is absent at compile time (dsu = b = 0: same bits as the general law with
// dsu = 0, without its registers -- carrying it at run time cost the linear-hardening case 14-30 %, profiles/r01l).
struct HosTrial {
  double st[6], s[6], pm, seq;
};

// trial stress, its deviator and the von Mises equivalent (same expressions as the J2 update)
DXM_HD void hos_trial(const double lam, const double mu, const double (&eps)[6], const double (&e_old)[6],
                      const double (&s_old)[6], HosTrial& t) {
  const double twomu = 2.0 * mu;
  double de[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) de[i] = eps[i] - e_old[i];
  const double tr = (de[0] + de[1]) + de[2];
  const double ltr = lam * tr;
#pragma unroll
  for (int i = 0; i < 3; ++i) t.st[i] = s_old[i] + fma_c(twomu, de[i], ltr);
#pragma unroll
  for (int i = 3; i < 6; ++i) t.st[i] = fma_c(twomu, de[i], s_old[i]);
  t.pm = ((t.st[0] + t.st[1]) + t.st[2]) / 3.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) t.s[i] = t.st[i] - t.pm;
#pragma unroll
  for (int i = 3; i < 6; ++i) t.s[i] = t.st[i];
  double ss = t.s[0] * t.s[0];
#pragma unroll
  for (int i = 1; i < 6; ++i) ss = fma_c(t.s[i], t.s[i], ss);
  t.seq = sqrt(1.5 * ss);
}

struct HosSol {
  HosRes cur;  // residual / criterion data at the solution (the tangent needs e.n, e.h, e.u, e.iphi, dsy)
  double dp, resid;
  int n_iter;
  bool flag, fail;
};

// 4-unknown Newton (3 principal deviatoric stresses + dp) in the principal axes of the trial deviator, eigenvalues l
template <int AT>
DXM_HD void hos_newton(const double (&l)[3], const double mu, const HosHard& hd, const double sy0, const double dsy0,
                       const double seq, const int a_rt, HosSol& o) {
  const int a = AT > 0 ? AT : a_rt;
  const double twomu = 2.0 * mu;
  const double threemu = 3.0 * mu;
  const double am1 = (double)a - 1.0, inv_a = 1.0 / (double)a;
  const double tol = kNewtonRtol * seq;
  o.flag = false;
  o.fail = false;
  o.n_iter = 0;
  o.resid = 0.0;
  double dp = 0.0;
  double x[3] = {0.0, 0.0, 0.0}, xe[3], dx[3] = {0.0, 0.0, 0.0}, dpe = 0.0, ddp = 0.0, t = 1.0;
  int ls = 0, stage = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) xe[k] = l[k];
  // One evaluation site of the criterion for its two uses (trial state, line-search candidates of a Newton step):
  // stage 0 = yield check at the trial state, 1 = start point, 2 = line-search candidate.  The start point is the trial
  // state scaled radially onto the yield surface: the criterion is homogeneous of degree one, so its data there follow
  // from the trial evaluation (phi scales, n / h / u do not change) -- no second evaluation.
  double m_prev = 0.0;
  for (;;) {
    // evaluated in place: once the step (dx, ddp) is formed only the merit value of the previous iterate is needed
    if (stage != 1) hos_eval(xe, a, inv_a, o.cur.e);
    hos_residual_finish(xe, dpe, l, twomu, hd, o.cur);
    if (stage == 0) {
      const double f = o.cur.e.phi - sy0;
      o.flag = f > 0.0;
      if (!o.flag) break;
      // start on the yield surface along the trial direction, dp from the J2-like estimate
      dpe = f / (threemu + dsy0);
      double sy1, dsy1;
      hos_hard(hd, dpe, sy1, dsy1);
      const double sc = sy1 / o.cur.e.phi;
#pragma unroll
      for (int k = 0; k < 3; ++k) xe[k] = l[k] * sc;
      o.cur.e.phi = o.cur.e.phi * sc;
      o.cur.e.iphi = o.cur.e.iphi / sc;
      stage = 1;
      continue;
    }
    if (stage == 2) {
      if (!(o.cur.m2 < m_prev || ls == kHosfordLsMax)) {  // no decrease: halve the step
        t = 0.5 * t;
        ++ls;
#pragma unroll
        for (int k = 0; k < 3; ++k) xe[k] = fma_c(t, dx[k], x[k]);
        dpe = fma_c(t, ddp, dp);
        continue;
      }
      ++o.n_iter;
    }
    stage = 2;
#pragma unroll
    for (int k = 0; k < 3; ++k) x[k] = xe[k];
    dp = dpe;
    m_prev = o.cur.m2;
    const double res = fmax(fmax(fabs(o.cur.rs[0]), fabs(o.cur.rs[1])), fmax(fabs(o.cur.rs[2]), fabs(o.cur.r4)));
    if (res <= tol) {
      o.resid = res;
      break;
    }
    if (o.n_iter == kNewtonCap || !(res == res)) {
      o.resid = res;
      o.fail = true;
      break;
    }
    // Schur complement on the adjugate: with y = adj(A) rs, z = adj(A) n (NOT divided by det A) the step in dp is
    // (r4 det - n.y) / (2 mu n.z + dsy det) -- one division on the critical path, 1/det runs beside it
    double Cf[6], det, y[3], z[3];
    hos_system(o.cur.e, twomu * dp, am1 * o.cur.e.iphi, Cf, det);
    hos_apply(Cf, o.cur.rs, y);
    hos_apply(Cf, o.cur.e.n, z);
    const double idet = 1.0 / det;
    const double ny = fma_c(o.cur.e.n[2], y[2], fma_c(o.cur.e.n[1], y[1], o.cur.e.n[0] * y[0]));
    const double nz = fma_c(o.cur.e.n[2], z[2], fma_c(o.cur.e.n[1], z[1], o.cur.e.n[0] * z[0]));
    ddp = fms_c(o.cur.r4, det, ny) / fma_c(twomu, nz, o.cur.dsy * det);
    const double tz = twomu * ddp;
#pragma unroll
    for (int k = 0; k < 3; ++k) dx[k] = -(fma_c(tz, z[k], y[k]) * idet);
    t = 1.0;
    ls = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) xe[k] = fma_c(t, dx[k], x[k]);
    dpe = fma_c(t, ddp, dp);
  }
  o.dp = dp;
}

// new state, stress and the 21 unique tangent entries (j <= i, row-major upper triangle = sym6_packed order) from the
// trial state, the old plastic strain and -- plastic points only -- the local solution and the eigenvectors V
template <int AT>
DXM_HD void hos_finish(const double lam, const double mu, const HosTrial& tr, const double p_old,
                       const double (&ep_old)[6], const bool plastic, const HosSol& so, const double (&V)[3][3],
                       const int a_rt, double (&sig)[6], double& p_new, double (&epsp)[6], double (&ct21)[21],
                       bool& fail) {
  const int a = AT > 0 ? AT : a_rt;
  const double twomu = 2.0 * mu;
  const HosRes& cur = so.cur;
  double dp = so.dp;
  double mN0[6], mN1[6], mN2[6], nrm[6];
  if (plastic) {
    hos_mandel_pair<0, 0>(V, mN0);
    hos_mandel_pair<1, 1>(V, mN1);
    hos_mandel_pair<2, 2>(V, mN2);
#pragma unroll
    for (int i = 0; i < 6; ++i) nrm[i] = fma_c(cur.e.n[2], mN2[i], fma_c(cur.e.n[1], mN1[i], cur.e.n[0] * mN0[i]));
  } else {
#pragma unroll
    for (int i = 0; i < 6; ++i) nrm[i] = 0.0;
    dp = 0.0;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const double depsp = dp * nrm[i];
    sig[i] = fnma_c(twomu, depsp, tr.st[i]);
    epsp[i] = ep_old[i] + depsp;
  }
  p_new = p_old + dp;

  if (!plastic) {
    const double AB = lam + twomu;
#pragma unroll
    for (int j = 0; j < 6; ++j)
#pragma unroll
      for (int i = j; i < 6; ++i)
        ct21[sym6_packed(j * 6 + i)] = (i == j) ? ((i < 3) ? AB : twomu) : ((i < 3 && j < 3) ? lam : 0.0);
  } else {
    double Cf[6], det, z[3];
    const double c = twomu * dp, iphi = cur.e.iphi;
    hos_system(cur.e, c, ((double)a - 1.0) * cur.e.iphi, Cf, det);
    hos_apply(Cf, cur.e.n, z);
    const double idet = 1.0 / det;
#pragma unroll
    for (int k = 0; k < 3; ++k) z[k] = z[k] * idet;
    const double nz = fma_c(cur.e.n[2], z[2], fma_c(cur.e.n[1], z[1], cur.e.n[0] * z[0]));
    const double w = (twomu * twomu) / fma_c(twomu, nz, cur.dsy);
    const double ti = twomu * idet;
    const double An00 = fnma_c(w, z[0] * z[0], fma_c(ti, Cf[0], lam));
    const double An01 = fnma_c(w, z[0] * z[1], fma_c(ti, Cf[1], lam));
    const double An02 = fnma_c(w, z[0] * z[2], fma_c(ti, Cf[2], lam));
    const double An11 = fnma_c(w, z[1] * z[1], fma_c(ti, Cf[3], lam));
    const double An12 = fnma_c(w, z[1] * z[2], fma_c(ti, Cf[4], lam));
    const double An22 = fnma_c(w, z[2] * z[2], fma_c(ti, Cf[5], lam));
    const double th01 = fma_c(0.5, hos_divdiff(-cur.e.u[2], cur.e.u[1], a), cur.e.h[0]) * iphi;
    const double th12 = fma_c(0.5, hos_divdiff(-cur.e.u[0], cur.e.u[2], a), cur.e.h[1]) * iphi;
    const double th20 = fma_c(0.5, hos_divdiff(-cur.e.u[1], cur.e.u[0], a), cur.e.h[2]) * iphi;
    const double G0 = twomu / fma_c(c, th01, 1.0), G1 = twomu / fma_c(c, th12, 1.0), G2 = twomu / fma_c(c, th20, 1.0);
    double mS0[6], mS1[6], mS2[6];
    hos_mandel_pair<0, 1>(V, mS0);
    hos_mandel_pair<1, 2>(V, mS1);
    hos_mandel_pair<2, 0>(V, mS2);
    double wN0[6], wN1[6], wN2[6];
#pragma unroll
    for (int cc = 0; cc < 6; ++cc) {
      wN0[cc] = fma_c(An02, mN2[cc], fma_c(An01, mN1[cc], An00 * mN0[cc]));
      wN1[cc] = fma_c(An12, mN2[cc], fma_c(An11, mN1[cc], An01 * mN0[cc]));
      wN2[cc] = fma_c(An22, mN2[cc], fma_c(An12, mN1[cc], An02 * mN0[cc]));
    }
    // the tangent is Q^T D Q with Q = rows (mN, mS) and D = blockdiag(An, diag G): wN = An mN above, gS = G mS here
    double gS0[6], gS1[6], gS2[6];
#pragma unroll
    for (int cc = 0; cc < 6; ++cc) {
      gS0[cc] = G0 * mS0[cc];
      gS1[cc] = G1 * mS1[cc];
      gS2[cc] = G2 * mS2[cc];
    }
#pragma unroll
    for (int j = 0; j < 6; ++j)
#pragma unroll
      for (int i = j; i < 6; ++i) {
        const double vs = fma_c(mS2[j], gS2[i], fma_c(mS1[j], gS1[i], mS0[j] * gS0[i]));
        ct21[sym6_packed(j * 6 + i)] = fma_c(mN2[j], wN2[i], fma_c(mN1[j], wN1[i], fma_c(mN0[j], wN0[i], vs)));
      }
  }
  double chk = (tr.seq + fabs(tr.pm)) + p_new;
#pragma unroll
  for (int i = 0; i < 6; ++i) chk = chk + fabs(epsp[i]);
  if (!isfinite(chk)) fail = true;
}

DXM_HD HosHard hos_hardening(const double sig0, const double H, const double dsu_rt, const double b_rt,
                             const double p_old, const bool voce) {
  const double dsu = voce ? dsu_rt : 0.0, b = voce ? b_rt : 0.0;
  return HosHard{sig0, H, dsu, b, voce ? b * dsu : 0.0, p_old};
}

// One Gauss point, the four stages back to back.  LIGHT == true: only the clearly elastic points are finished; for a
// candidate (the cheap rejection did not fire) the routine returns true without touching the outputs and the caller
// hands the point to the full routine.
template <bool LIGHT, int AT, bool VOCE>
DXM_HD bool hosford_point(const double lam, const double mu, const double sig0, const double H, const double dsu_rt,
                          const double b_rt, const int a_rt, const double bound, const double (&eps)[6], const double (&e_old)[6], const double (&s_old)[6],
                          const double p_old, const double (&ep_old)[6], double (&sig)[6], double& p_new,
                          double (&epsp)[6], double (&ct21)[21], bool& flag, int& n_iter, double& resid,
                          bool& fail) {
  HosTrial tr;
  hos_trial(lam, mu, eps, e_old, s_old, tr);
  const HosHard hd = hos_hardening(sig0, H, dsu_rt, b_rt, p_old, VOCE);
  double sy0, dsy0;
  hos_hard(hd, 0.0, sy0, dsy0);
  // sigma_eq <= bound * seq_Mises: below that the point is surely elastic
  const bool candidate = bound * tr.seq > sy0;
  if (LIGHT && candidate) return true;
  HosSol so;
  so.flag = false;
  so.fail = false;
  so.n_iter = 0;
  so.resid = 0.0;
  so.dp = 0.0;
  double V[3][3];
  if (!LIGHT && candidate) {
    double l[3];
    hos_eig3(tr.s, l, V);
    hos_newton<AT>(l, mu, hd, sy0, dsy0, tr.seq, a_rt, so);
  }
  flag = so.flag;
  n_iter = so.n_iter;
  resid = so.resid;
  fail = so.fail;
  hos_finish<AT>(lam, mu, tr, p_old, ep_old, !LIGHT && so.flag, so, V, a_rt, sig, p_new, epsp, ct21, fail);
  return false;
}

#if defined(__CUDACC__) && defined(DXM_HOSFORD_KERNELS)  // kernels: instantiated in dxm_hosford_api.cu only
// SmallStrainArgs is shared with the J2 kernels (same SoA state layout and hardening law); a.hos_a = exponent,
// a.hos_bound = candidate bound.  Per-point
// properties (a.pE != nullptr) and diagnostics (a.d_flag != nullptr) are run-time switches here: the kernels are
// templated on the exponent only.
struct HosPointIO {
  double eps[6], e_old[6], s_old[6], ep_old[6], p_old, lam, mu, sig0, H, dsu, b;
};

// STREAM: evict-first loads (last use of the inputs); the first pass over a point keeps them cacheable for the second
template <bool STREAM>
__device__ __forceinline__ double hos_ld(const double* p) { return STREAM ? __ldcs(p) : __ldg(p); }

// the three 6-vectors the trial stress is formed from
template <bool STREAM>
__device__ __forceinline__ void hos_load_trial(const SmallStrainArgs& a, int64_t i0, HosPointIO& io) {
  const int64_t ld = a.ld;
#pragma unroll
  for (int c = 0; c < 6; ++c) io.eps[c] = hos_ld<STREAM>(a.eps + c * ld + i0);
#pragma unroll
  for (int c = 0; c < 6; ++c) io.e_old[c] = hos_ld<STREAM>(a.eps_old + c * ld + i0);
#pragma unroll
  for (int c = 0; c < 6; ++c) io.s_old[c] = hos_ld<STREAM>(a.sig_old + c * ld + i0);
}

template <bool STREAM, bool VOCE>
__device__ __forceinline__ void hos_load_props(const SmallStrainArgs& a, int64_t i0, HosPointIO& io) {
  io.p_old = hos_ld<STREAM>(a.p_old + i0);
  io.lam = a.lam;
  io.mu = a.mu;
  io.sig0 = a.sig0;
  io.H = a.H;
  io.dsu = a.dsu;
  io.b = a.b;
  if (a.pE) {
    const double E = hos_ld<STREAM>(a.pE + i0), nu = hos_ld<STREAM>(a.pnu + i0);
    io.lam = E * nu / (1.0 + nu) / (1.0 - 2.0 * nu);
    io.mu = E / 2.0 / (1.0 + nu);
    io.sig0 = hos_ld<STREAM>(a.psig0 + i0);
    io.H = hos_ld<STREAM>(a.pH + i0);
    if (VOCE) {
      const double d = hos_ld<STREAM>(a.psigu + i0) - io.sig0;
      io.dsu = isfinite(d) ? d : 0.0;
      io.b = hos_ld<STREAM>(a.pb + i0);
    }
  }
}

template <bool STREAM, bool VOCE>
__device__ __forceinline__ void hos_load(const SmallStrainArgs& a, int64_t i0, HosPointIO& io) {
  hos_load_trial<STREAM>(a, i0, io);
  hos_load_props<STREAM, VOCE>(a, i0, io);
#pragma unroll
  for (int c = 0; c < 6; ++c) io.ep_old[c] = hos_ld<STREAM>(a.epsp_old + c * a.ld + i0);
}

__device__ __forceinline__ void hos_finish_store(const SmallStrainArgs& a, int64_t i0, const double (&sig)[6], double p_new,
                                           const double (&epsp)[6], const double (&ct21)[21], bool flag, int n_iter,
                                           double resid, bool fail, PointStats& acc) {
  const int64_t ld = a.ld;
  acc.n_plastic += flag ? 1u : 0u;
  acc.n_fail += fail ? 1u : 0u;
  acc.max_iter = n_iter > (int)acc.max_iter ? (unsigned)n_iter : acc.max_iter;
  acc.max_resid = resid > acc.max_resid ? resid : acc.max_resid;
  if (resid != resid) acc.max_resid = resid;
  if (a.d_flag) {
    a.d_flag[i0] = flag ? 1 : 0;
    a.d_iter[i0] = n_iter;
    a.d_resid[i0] = resid;
    a.d_fail[i0] = fail ? 1 : 0;
  }
#pragma unroll
  for (int c = 0; c < 6; ++c) __stcs(a.sig + c * ld + i0, sig[c]);
  __stcs(a.p + i0, p_new);
#pragma unroll
  for (int c = 0; c < 6; ++c) __stcs(a.epsp + c * ld + i0, epsp[c]);
#pragma unroll
  for (int r = 0; r < 21; ++r) __stcs(a.ct + (int64_t)r * ld + i0, ct21[r]);
}

constexpr int kHosBlock = 128;
struct HosPark {
  double V[9][kHosBlock];  // eigenvectors of the thread's point while its Newton loop runs
};

// Full update of the point at SoA position i0 by the calling thread, register-lean: across the Newton loop only the
// eigenvalues, the loop state and the hardening constants stay in registers -- the eigenvectors wait in shared memory
// and the trial stress is formed a second time from the inputs (cacheable first pass -> L2 hits) once the loop is over.
// Same stages, same operations on the same values as hosford_point(): bit-identical results.
template <int AT, bool VOCE>
__device__ __forceinline__ void hos_solve_point(const SmallStrainArgs& a, const int64_t i0, HosPark& park, PointStats& acc) {
  HosPointIO io;
  hos_load_props<false, VOCE>(a, i0, io);
  const HosHard hd = hos_hardening(io.sig0, io.H, io.dsu, io.b, io.p_old, VOCE);
  double sy0, dsy0;
  hos_hard(hd, 0.0, sy0, dsy0);
  HosSol so;
  so.flag = false;
  so.fail = false;
  so.n_iter = 0;
  so.resid = 0.0;
  so.dp = 0.0;
  {
    HosTrial tr;
    hos_load_trial<false>(a, i0, io);
    hos_trial(io.lam, io.mu, io.eps, io.e_old, io.s_old, tr);
    if (a.hos_bound * tr.seq > sy0) {
      double l[3], V[3][3];
      hos_eig3(tr.s, l, V);
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) park.V[r * 3 + c][threadIdx.x] = V[r][c];
      hos_newton<AT>(l, io.mu, hd, sy0, dsy0, tr.seq, a.hos_a, so);
    }
  }
  HosTrial tr;
  hos_load_trial<true>(a, i0, io);
  hos_trial(io.lam, io.mu, io.eps, io.e_old, io.s_old, tr);
#pragma unroll
  for (int c = 0; c < 6; ++c) io.ep_old[c] = __ldcs(a.epsp_old + c * a.ld + i0);
  double V[3][3];
  if (so.flag) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) V[r][c] = park.V[r * 3 + c][threadIdx.x];
  }
  double sig[6], epsp[6], ct21[21], p_new;
  bool fail = so.fail;
  hos_finish<AT>(io.lam, io.mu, tr, io.p_old, io.ep_old, so.flag, so, V, a.hos_a, sig, p_new, epsp, ct21, fail);
  hos_finish_store(a, i0, sig, p_new, epsp, ct21, so.flag, so.n_iter, so.resid, fail, acc);
}

// fused: every thread runs the full routine on its own point (small batches, mostly-plastic batches, A/B reference)
template <int AT, bool VOCE, int MINB>
__global__ void __launch_bounds__(kHosBlock, MINB) dxm_hosford_kernel(const SmallStrainArgs a) {
  __shared__ HosPark park;
  const int64_t ntile = (a.count + blockDim.x - 1) / blockDim.x;
  PointStats acc;
  for (int64_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int64_t loc = tile * blockDim.x + threadIdx.x;
    if (loc >= a.count) continue;
    hos_solve_point<AT, VOCE>(a, a.start + loc, park, acc);
  }
  block_reduce_stats(acc, a.stats);
}

// tiled: one CTA owns a tile of kHosTile consecutive points.  Phase A streams the tile (clearly elastic points are
// finished, candidates go to a shared-memory queue with one warp-aggregated atomic per warp, order kept); phase B
// packs the candidates into full warps for the local solves.  The candidates' stores land next to the neighbours'
// stores issued moments earlier by the same CTA -- a device-wide queue + second kernel (r01g) lost exactly that
// locality when few points are candidates (scattered 8-byte accesses) and was slower in every regime.
constexpr int kHosTile = 1024;
template <int AT, bool VOCE, int MINB>
__global__ void __launch_bounds__(kHosBlock, MINB) dxm_hosford_tiled_kernel(const SmallStrainArgs a) {
  __shared__ HosPark park;
  __shared__ unsigned s_queue[kHosTile];
  __shared__ unsigned s_count;
  const int64_t ntile = (a.count + kHosTile - 1) / kHosTile;
  PointStats acc;
  for (int64_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    for (int sub = 0; sub < kHosTile / kHosBlock; ++sub) {
      const int64_t loc = tile * kHosTile + sub * kHosBlock + threadIdx.x;
      bool heavy = false;
      if (loc < a.count) {
        const int64_t i0 = a.start + loc;
        HosPointIO io;
        hos_load<false, VOCE>(a, i0, io);
        double sig[6], epsp[6], ct21[21], p_new, resid;
        bool flag, fail;
        int n_iter;
        heavy = hosford_point<true, 0, VOCE>(io.lam, io.mu, io.sig0, io.H, io.dsu, io.b, a.hos_a, a.hos_bound, io.eps, io.e_old, io.s_old,
                                       io.p_old, io.ep_old, sig, p_new, epsp, ct21, flag, n_iter, resid, fail);
        if (!heavy) hos_finish_store(a, i0, sig, p_new, epsp, ct21, flag, n_iter, resid, fail, acc);
      }
      const unsigned bal = __ballot_sync(0xffffffffu, heavy);
      if (bal) {
        const int lane = threadIdx.x & 31;
        unsigned base = 0;
        if (lane == __ffs(bal) - 1) base = atomicAdd(&s_count, (unsigned)__popc(bal));
        base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
        if (heavy) s_queue[base + __popc(bal & ((1u << lane) - 1u))] = (unsigned)(loc - tile * kHosTile);
      }
    }
    __syncthreads();
    const unsigned total = s_count;
    for (unsigned q = threadIdx.x; q < total; q += blockDim.x)
      hos_solve_point<AT, VOCE>(a, a.start + tile * kHosTile + (int64_t)s_queue[q], park, acc);
    __syncthreads();  // the queue is reused by the next tile
  }
  block_reduce_stats(acc, a.stats);
}

#endif  // kernels

#ifdef __CUDACC__
// host side of the Hosford launches (dxm_hosford_api.cu)
struct HosLaunch {
  int num_sms;
  cudaStream_t stream;
  bool voce;   // the hardening law has a saturation term (sigu was set): general-law instantiation
  bool tiled;  // tiled kernel (stream + CTA-local candidate queue + packed local solves) instead of the fused one
  int tiles_per_cta;
  int minb;  // resident CTAs per SM the register allocation targets: 4 (128 registers) or 3 (168); 0 = per-kernel default
};
int launch_hosford(const SmallStrainArgs& a, const HosLaunch& cfg, int* launches);
double hosford_bound(int a);
#endif  // __CUDACC__

#undef DXM_HD
}  // namespace dxm
