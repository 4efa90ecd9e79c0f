/*
 * dxm_oracle_hosford.c -- plain-C oracle of the small-strain Hosford plasticity update.  TEST INFRASTRUCTURE ONLY
 * (see oracle/__init__.py); never linked into the product.
 *
 * Restates the behaviour of demos/multimaterials/IsotropicPlasticHosfordFlowLinear.mfront:1-27 (MFront `Implicit` DSL,
 * StandardElastoViscoPlasticity brick: Hooke stress potential, "Plastic" flow with the Hosford criterion {a: 10} and
 * linear isotropic hardening R0 + H p, theta = 1), the matrix phase of demos/multimaterials/multimaterials.py:245-254
 * (generalised to the hardening law of the J2 behaviours, sig0 + H p + (sigu - sig0)(1 - exp(-b p)), i.e. jaxmat's
 * GeneralIsotropicHardening(elastic_model, yield_stress, ...) of demos/jax/elastoplasticity/_plane_stress_elastoplasticity.py:45):
 *
 *     sigma_eq = ( 1/2 (|s1-s2|^a + |s2-s3|^a + |s3-s1|^a) )^(1/a)            principal stresses s_k
 *     eel + dp n(sigma) = eel_old + deps ,  n = d sigma_eq / d sigma          (backward Euler, associated flow)
 *     sigma_eq(sigma) - R0 - H (p_old + dp) = 0 ,  sigma = C : eel
 *
 * MFront's generated code is not in the tree and TFEL/MGIS are absent, so parity with MFront is UNPINNED; the
 * restatement is checked against an independent solve of the 7-unknown system above (numpy eigh + scipy) and finite
 * differences (tests/test_oracle_hosford.py).  State layout follows the other small-strain behaviours of this build
 * (strain, stress, p, epsp; eel = strain - epsp); the exponent a is an even integer (2 = von Mises, 6/8 the usual
 * bcc/fcc fits, 10 the demo), which makes every power a product chain and the tangent's spectral terms exact.
 *
 * Algorithm (operation order == csrc/dxm_hosford.cuh, compiled with -ffp-contract=off / -fmad=false):
 *   trial stress as in the J2 update; cheap rejection sigma_eq <= (2^(a-1)+1)^(1/a)/sqrt(3) seq_Mises (pure shear);
 *   non-iterative eigen-decomposition of the trial deviator (eig3 below; +, -, *, /, sqrt only; the equivalent stress
 *   costs one division per evaluation: the a-th root is a division-free fixed-count iteration on q^(-1/a)); isotropy keeps the principal
 *   axes, so the return map is a 4-unknown Newton (3 principal deviatoric stresses + dp) started from the radially
 *   scaled trial state with a simple-decrease backtracking line search; the consistent tangent
 *   Xi - (Xi n)(Xi n)^T / (n Xi n + H), Xi = (C^-1 + dp dn/dsigma)^-1, is assembled from its spectral form: a 3x3
 *   normal block 2 mu A^-1 + lam 1 1^T - ..., and three shear moduli 2 mu / (1 + 2 mu dp theta_ij) with
 *   theta_ij = (n_i - n_j)/(s_i - s_j) written as an exact divided difference (no 0/0 at repeated eigenvalues).
 */
#include "dxm_canon.h"

#define RSQRT2 0.7071067811865476
#define SQRT2 1.4142135623730951
#define LS_MAX 10

/* isotropic hardening sigma_Y(p) = sig0 + H p + dsu (1 - exp(-b p)) (the law of the J2 behaviours: linear for
 * dsu = 0, Voce for H = 0) and its slope, at p = p_old + dp */
typedef struct {
  double sig0, H, dsu, b, bdsu, p_old;
} hard_t;

static void hard_eval(const hard_t* hd, double dp, double* sy, double* dsy) {
  const double p = hd->p_old + dp;
  const double e = (hd->bdsu != 0.0) ? exp_c(-(hd->b * p)) : 1.0;
  *sy = FMA(hd->dsu, 1.0 - e, FMA(hd->H, p, hd->sig0));
  *dsy = FMA(hd->bdsu, e, hd->H);
}

/* (x*x)^k, 1 <= k <= 32, by binary powering from the top bit of k: y = x^2; per lower bit: y = y*y, then y = y * x^2 if
 * the bit is set (k = 5: x^2, x^4, x^8, x^10) */
static double ipow2(double x, int k) {
  const double x2 = x * x;
  double y = x2;
  int top = 5;
  while (top > 0 && !((k >> top) & 1)) --top;
  for (int bit = 4; bit >= 0; --bit) {
    if (bit < top) {
      y = y * y;
      if ((k >> bit) & 1) y = y * x2;
    }
  }
  return y;
}

/* q^(-1/a), q in (0.5, 1], division free: second-order Taylor start in x = 1 - q, then a fixed two steps of the
 * third-order correction: r = 1 - q w^a, exact root = w (1 - r)^(-1/a) = w (1 + s r (1 + (s+1)/2 r (1 + (s+2)/3 r ...))),
 * s = 1/a, truncated after r^3 (error e -> O(e^4): below 2e-18 after two steps for every even a in [2, 64]) */
#define ROOT_STEPS 2
static double arootinv(double q, int a, double inv_a) {
  const double x = 1.0 - q;
  const double k2 = 0.5 * (1.0 + inv_a);
  const double k3 = (2.0 + inv_a) / 3.0;
  double w = FMA(x * inv_a, FMA(x, k2, 1.0), 1.0);
  for (int it = 0; it < ROOT_STEPS; ++it) {
    const double r = FNMA(q, ipow2(w, a / 2), 1.0);
    w = w * FMA(r * inv_a, FMA(r * k2, FMA(r, k3, 1.0), 1.0), 1.0);
  }
  return w;
}

/* the root alone, for the accuracy scan of tests/test_oracle_hosford.py */
void dxo_hosford_root(int64_t n, const double* q, int a, double* w) {
  for (int64_t i = 0; i < n; ++i) w[i] = arootinv(q[i], a, 1.0 / (double)a);
}

/* Hosford equivalent stress of principal values l, 1/phi, its gradient n, h_k = (d_k/phi)^(a-2), u_k = d_k/phi.
 * One division: 1/max|d|. */
static void hosford_eval(const double l[3], int a, double inv_a, double* phi, double* iphi, double n[3], double h[3],
                         double u[3]) {
  const double d0 = l[0] - l[1], d1 = l[1] - l[2], d2 = l[2] - l[0];
  const double m = fmax(fmax(fabs(d0), fabs(d1)), fabs(d2));
  const double im = 1.0 / m;
  const double r0 = d0 * im, r1 = d1 * im, r2 = d2 * im;
  const double q = 0.5 * ((ipow2(r0, a / 2) + ipow2(r1, a / 2)) + ipow2(r2, a / 2));
  const double w = arootinv(q, a, inv_a);
  *phi = m * (q * ((a > 2) ? ipow2(w, (a - 2) / 2) * w : w)); /* m q^(1/a) = m q w^(a-1): no division */
  *iphi = w * im;
  u[0] = r0 * w;
  u[1] = r1 * w;
  u[2] = r2 * w;
  for (int k = 0; k < 3; ++k) h[k] = (a > 2) ? ipow2(u[k], (a - 2) / 2) : 1.0;
  const double g0 = h[0] * u[0], g1 = h[1] * u[1], g2 = h[2] * u[2];
  n[0] = 0.5 * (g0 - g2);
  n[1] = 0.5 * (g1 - g0);
  n[2] = 0.5 * (g2 - g1);
}

/* sum_{k=0}^{a-2} x^k y^(a-2-k)  ( = (x^(a-1) - y^(a-1)) / (x - y), exact also at x == y ) */
static double divdiff(double x, double y, int a) {
  double t = 1.0, xp = 1.0;
  for (int j = 1; j <= a - 2; ++j) {
    xp = xp * x;
    t = FMA(y, t, xp);
  }
  return t;
}

/* ---- symmetric 3x3 eigen-decomposition of a deviator, non-iterative ----------------------------------------------
 * (replaces the cyclic Jacobi sweeps of round 1 / early round 2: 4-5 sweeps x 3 rotations, each with two square roots and
 * two divisions -- a fifth of the update's time on the GPU.)  With B = A / p, p = sqrt(tr(A^2) / 6), the eigenvalues of B
 * are 2 cos(theta + 2 pi j / 3), cos(3 theta) = det(B) / 2 =: k.  The eigenvalue on the side of the sign of k is ISOLATED
 * (at least 0.866 * 2 away from the other two whatever the spectrum), so
 *   1. y = cos(theta) in [sqrt(3)/2, 1] solves 4 y^3 - 3 y = |k|: cubic start polynomial + three division-free Newton
 *      steps (the reciprocal slope R is refined alongside: R <- R (2 - g' R)); beta0 = sgn(k) 2 y;
 *   2. its eigenvector n = the largest of the three cross products of rows of B - beta0 I, normalised (well conditioned
 *      because beta0 is isolated);
 *   3. an orthonormal basis (U, W) of the plane normal to n without a square root (Duff et al., "Building an orthonormal
 *      basis, revisited", JCGT 2017), and ONE exact Jacobi rotation of the 2x2 restriction of B to that plane -- repeated
 *      or nearly repeated eigenvalues there are harmless (any rotation of an eigenplane is a valid basis).
 * Eigenvalues are Rayleigh quotients of the computed vectors.  Residual |A V - V L| / |A| and |V^T V - I| <= ~1e-15 over
 * random, axisymmetric, nearly degenerate and pure-shear spectra (tests/test_oracle_hosford.py). */
#define EIG_CY0 0.8660615980506479
#define EIG_CY1 0.16540585304875938
#define EIG_CY2 -0.04088323804684024
#define EIG_CY3 0.009444179663245256
#define EIG_CR0 0.1660512323983123
#define EIG_CR1 -0.08345496691468178
#define EIG_CR2 0.028968785247117986
#define EIG_SIXTH 0.16666666666666666

static void cross3(const double u[3], const double v[3], double c[3], double* d) {
  c[0] = FMS(u[1], v[2], u[2] * v[1]);
  c[1] = FMS(u[2], v[0], u[0] * v[2]);
  c[2] = FMS(u[0], v[1], u[1] * v[0]);
  *d = FMA(c[2], c[2], FMA(c[1], c[1], c[0] * c[0]));
}

/* symmetric 3x3 from a Mandel deviator: eigenvalues l, eigenvectors in the columns of Q */
static void eig3(const double s[6], double l[3], double Q[3][3]) {
  const double a00 = s[0], a11 = s[1], a22 = s[2], a01 = s[3] * RSQRT2, a02 = s[4] * RSQRT2, a12 = s[5] * RSQRT2;
  const double dg = FMA(a22, a22, FMA(a11, a11, a00 * a00));
  const double od = FMA(a12, a12, FMA(a02, a02, a01 * a01));
  const double p2 = FMA(2.0, od, dg) * EIG_SIXTH;
  if (p2 == 0.0) { /* zero deviator (never a candidate point) */
    for (int i = 0; i < 3; ++i) {
      l[i] = 0.0;
      for (int j = 0; j < 3; ++j) Q[i][j] = (i == j) ? 1.0 : 0.0;
    }
    return;
  }
  const double p = sqrt(p2), ip = 1.0 / p;
  const double b00 = a00 * ip, b11 = a11 * ip, b22 = a22 * ip, b01 = a01 * ip, b02 = a02 * ip, b12 = a12 * ip;
  const double m0 = FMS(b11, b22, b12 * b12), m1 = FMS(b01, b22, b12 * b02), m2 = FMS(b01, b12, b11 * b02);
  const double hdet = 0.5 * FMA(b02, m2, FMS(b00, m0, b01 * m1));
  const double sgn = (hdet >= 0.0) ? 1.0 : -1.0;
  const double k = fmin(fabs(hdet), 1.0);
  double y = FMA(FMA(FMA(EIG_CY3, k, EIG_CY2), k, EIG_CY1), k, EIG_CY0);
  double R = FMA(FMA(EIG_CR2, k, EIG_CR1), k, EIG_CR0);
  for (int it = 0; it < 3; ++it) {
    const double y2 = y * y;
    const double g = FMS(FMS(4.0, y2, 3.0), y, k);
    const double gp = FMS(12.0, y2, 3.0);
    R = R * FNMA(gp, R, 2.0);
    y = FNMA(g, R, y);
  }
  const double beta = sgn * (2.0 * y);
  const double r0[3] = {b00 - beta, b01, b02}, r1[3] = {b01, b11 - beta, b12}, r2[3] = {b02, b12, b22 - beta};
  double c[3], d, c2[3], d2;
  cross3(r0, r1, c, &d);
  cross3(r0, r2, c2, &d2);
  if (d2 > d) { c[0] = c2[0]; c[1] = c2[1]; c[2] = c2[2]; d = d2; }
  cross3(r1, r2, c2, &d2);
  if (d2 > d) { c[0] = c2[0]; c[1] = c2[1]; c[2] = c2[2]; d = d2; }
  const double inv = 1.0 / sqrt(d);
  const double n[3] = {c[0] * inv, c[1] * inv, c[2] * inv};
  const double sg = (n[2] >= 0.0) ? 1.0 : -1.0;
  const double a = -1.0 / (sg + n[2]);
  const double bq = (n[0] * n[1]) * a;
  const double U[3] = {FMA(sg * (n[0] * n[0]), a, 1.0), sg * bq, -(sg * n[0])};
  const double W[3] = {bq, FMA(n[1] * n[1], a, sg), -n[1]};
#define EIG_MV(v, o)                                        \
  o[0] = FMA(b02, v[2], FMA(b01, v[1], b00 * v[0]));        \
  o[1] = FMA(b12, v[2], FMA(b11, v[1], b01 * v[0]));        \
  o[2] = FMA(b22, v[2], FMA(b12, v[1], b02 * v[0]));
  double Bn[3], BU[3], BW[3];
  EIG_MV(n, Bn)
  EIG_MV(U, BU)
  EIG_MV(W, BW)
#undef EIG_MV
  const double l0 = dot3(n[0], Bn[0], n[1], Bn[1], n[2], Bn[2]);
  const double m00 = dot3(U[0], BU[0], U[1], BU[1], U[2], BU[2]);
  const double m01 = dot3(U[0], BW[0], U[1], BW[1], U[2], BW[2]);
  const double m11 = dot3(W[0], BW[0], W[1], BW[1], W[2], BW[2]);
  /* the Jacobi rotation of [[m00, m01], [m01, m11]]: t = tan of the smaller angle, one division */
  const double delta = (m11 - m00) * 0.5;
  const double den = fabs(delta) + sqrt(FMA(delta, delta, m01 * m01));
  double t = (den > 0.0) ? m01 / den : 0.0;
  if (delta < 0.0) t = -t;
  const double cs = 1.0 / sqrt(FMA(t, t, 1.0)), sn = t * cs;
  l[0] = l0 * p;
  l[1] = FNMA(t, m01, m00) * p;
  l[2] = FMA(t, m01, m11) * p;
  for (int i = 0; i < 3; ++i) {
    Q[i][0] = n[i];
    Q[i][1] = FMS(cs, U[i], sn * W[i]);
    Q[i][2] = FMA(sn, U[i], cs * W[i]);
  }
}

/* the eigen-decomposition alone, for the accuracy scan of tests/test_oracle_hosford.py: s (n, 6) Mandel deviators ->
 * l (n, 3), Q (n, 3, 3) */
void dxo_hosford_eig(int64_t n, const double* s, double* l, double* Q) {
  for (int64_t i = 0; i < n; ++i) {
    double Qi[3][3];
    eig3(s + 6 * i, l + 3 * i, Qi);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) Q[9 * i + 3 * r + c] = Qi[r][c];
  }
}

typedef struct {
  double rs[3], r4, phi, iphi, n[3], h[3], u[3], m2, dsy;
} hres_t;

/* residuals of the 4 unknowns (x, dp) from the criterion data (phi, n, ...) already evaluated at x */
static void residual_finish(const double x[3], double dp, const double l[3], double twomu, const hard_t* hd, hres_t* o) {
  const double c = twomu * dp;
  for (int k = 0; k < 3; ++k) o->rs[k] = FMA(c, o->n[k], x[k] - l[k]);
  double sy;
  hard_eval(hd, dp, &sy, &o->dsy);
  o->r4 = o->phi - sy;
  o->m2 = FMA(o->r4, o->r4, FMA(o->rs[2], o->rs[2], FMA(o->rs[1], o->rs[1], o->rs[0] * o->rs[0])));
}

static void hosford_residual(const double x[3], double dp, const double l[3], double twomu, const hard_t* hd, int a,
                             double inv_a, hres_t* o) {
  hosford_eval(x, a, inv_a, &o->phi, &o->iphi, o->n, o->h, o->u);
  residual_finish(x, dp, l, twomu, hd, o);
}

/* A = I + c k1 (M/2 - n n^T) (symmetric), its adjugate C and det */
static void hosford_system(const hres_t* r, double c, double k1, double Cf[6], double* det) {
  const double ck = c * k1;
  const double A00 = FMA(ck, FNMA(r->n[0], r->n[0], 0.5 * (r->h[0] + r->h[2])), 1.0);
  const double A11 = FMA(ck, FNMA(r->n[1], r->n[1], 0.5 * (r->h[0] + r->h[1])), 1.0);
  const double A22 = FMA(ck, FNMA(r->n[2], r->n[2], 0.5 * (r->h[1] + r->h[2])), 1.0);
  const double A01 = ck * FNMA(r->n[0], r->n[1], -0.5 * r->h[0]);
  const double A02 = ck * FNMA(r->n[0], r->n[2], -0.5 * r->h[2]);
  const double A12 = ck * FNMA(r->n[1], r->n[2], -0.5 * r->h[1]);
  Cf[0] = FMS(A11, A22, A12 * A12); /* C00 */
  Cf[1] = FMS(A02, A12, A01 * A22); /* C01 */
  Cf[2] = FMS(A01, A12, A02 * A11); /* C02 */
  Cf[3] = FMS(A00, A22, A02 * A02); /* C11 */
  Cf[4] = FMS(A01, A02, A00 * A12); /* C12 */
  Cf[5] = FMS(A00, A11, A01 * A01); /* C22 */
  *det = FMA(A02, Cf[2], FMA(A01, Cf[1], A00 * Cf[0]));
}

/* adj(A) v: the caller scales by 1/det where it needs A^-1 v */
static void sym3_apply(const double Cf[6], const double v[3], double o[3]) {
  o[0] = FMA(Cf[2], v[2], FMA(Cf[1], v[1], Cf[0] * v[0]));
  o[1] = FMA(Cf[4], v[2], FMA(Cf[3], v[1], Cf[1] * v[0]));
  o[2] = FMA(Cf[5], v[2], FMA(Cf[4], v[1], Cf[2] * v[0]));
}

/* props: E, nu, sig0 (R0), H, sigu, b scalars or per point (pp/per as in dxm_oracle.c); a even integer >= 2 */
void dxo_hosford(int64_t n, const double* eps, const double* e_old, const double* s_old, const double* p_old_a,
                 const double* ep_old, const double* const* pp, const int* per, int a, int newton_cap, double rtol,
                 double bound,
                 double* sig_o, double* p_o, double* epsp_o, double* ct_o, uint8_t* flag_o, int32_t* iter_o,
                 double* resid_o, uint8_t* fail_o) {
  for (int64_t pt = 0; pt < n; ++pt) {
    const double E = per[0] ? pp[0][pt] : pp[0][0], nu = per[1] ? pp[1][pt] : pp[1][0];
    const double sig0 = per[2] ? pp[2][pt] : pp[2][0], H = per[3] ? pp[3][pt] : pp[3][0];
    const double sigu = per[4] ? pp[4][pt] : pp[4][0], bb = per[5] ? pp[5][pt] : pp[5][0];
    double dsu = sigu - sig0;
    if (!isfinite(dsu)) dsu = 0.0;
    const double lam = E * nu / (1 + nu) / (1 - 2 * nu);
    const double mu = E / 2 / (1 + nu);
    const double twomu = 2.0 * mu, threemu = 3.0 * mu;
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
    const hard_t hd = {sig0, H, dsu, bb, bb * dsu, p_old};
    double sy0, dsy0;
    hard_eval(&hd, 0.0, &sy0, &dsy0);
    const double am1 = (double)a - 1.0, inv_a = 1.0 / (double)a;

    int flag = 0, n_iter = 0, fail = 0;
    double dp = 0.0, resid = 0.0;
    double l[3], Q[3][3];
    hres_t cur;
    if (bound * seq > sy0) { /* sigma_eq <= bound * seq (bound >= (2^(a-1)+1)^(1/a)/sqrt(3)): otherwise surely elastic */
      eig3(s, l, Q);
      hosford_eval(l, a, inv_a, &cur.phi, &cur.iphi, cur.n, cur.h, cur.u);
      const double f = cur.phi - sy0;
      flag = f > 0.0;
      if (flag) {
        /* start on the yield surface along the trial direction, dp from the J2-like estimate */
        dp = f / (threemu + dsy0);
        double sy1, dsy1;
        hard_eval(&hd, dp, &sy1, &dsy1);
        const double sc = sy1 / cur.phi;
        double x[3] = {l[0] * sc, l[1] * sc, l[2] * sc};
        /* the criterion is homogeneous of degree one: its data at the radially scaled start point follow from the trial
         * evaluation (phi scales, n / h / u do not change) -- no second evaluation */
        cur.phi = cur.phi * sc;
        cur.iphi = cur.iphi / sc;
        residual_finish(x, dp, l, twomu, &hd, &cur);
        const double tol = rtol * seq;
        for (int it = 0;; ++it) {
          const double res = fmax(fmax(fabs(cur.rs[0]), fabs(cur.rs[1])), fmax(fabs(cur.rs[2]), fabs(cur.r4)));
          if (res <= tol) { resid = res; break; }
          if (it == newton_cap || !(res == res)) { resid = res; fail = 1; break; }
          /* Schur complement on the adjugate (y, z not divided by det A): one division on the dependent chain */
          double Cf[6], det, y[3], z[3];
          hosford_system(&cur, twomu * dp, am1 * cur.iphi, Cf, &det);
          sym3_apply(Cf, cur.rs, y);
          sym3_apply(Cf, cur.n, z);
          const double idet = 1.0 / det;
          const double ny = dot3(cur.n[0], y[0], cur.n[1], y[1], cur.n[2], y[2]);
          const double nz = dot3(cur.n[0], z[0], cur.n[1], z[1], cur.n[2], z[2]);
          const double ddp = FMS(cur.r4, det, ny) / FMA(twomu, nz, cur.dsy * det);
          const double tz = twomu * ddp;
          const double dx[3] = {-(FMA(tz, z[0], y[0]) * idet), -(FMA(tz, z[1], y[1]) * idet),
                                -(FMA(tz, z[2], y[2]) * idet)};
          double t = 1.0;
          hres_t nxt;
          double xn[3], dpn;
          for (int ls = 0;; ++ls) {
            for (int k = 0; k < 3; ++k) xn[k] = FMA(t, dx[k], x[k]);
            dpn = FMA(t, ddp, dp);
            hosford_residual(xn, dpn, l, twomu, &hd, a, inv_a, &nxt);
            if (nxt.m2 < cur.m2 || ls == LS_MAX) break;
            t = 0.5 * t;
          }
          for (int k = 0; k < 3; ++k) x[k] = xn[k];
          dp = dpn;
          cur = nxt;
          ++n_iter;
        }
      }
    }

    /* flow direction in the global Mandel basis, state update */
    double mN[3][6], nrm[6];
    if (flag) {
      for (int k = 0; k < 3; ++k) {
        mN[k][0] = Q[0][k] * Q[0][k];
        mN[k][1] = Q[1][k] * Q[1][k];
        mN[k][2] = Q[2][k] * Q[2][k];
        mN[k][3] = SQRT2 * (Q[0][k] * Q[1][k]);
        mN[k][4] = SQRT2 * (Q[0][k] * Q[2][k]);
        mN[k][5] = SQRT2 * (Q[1][k] * Q[2][k]);
      }
      for (int i = 0; i < 6; ++i) nrm[i] = dot3(cur.n[0], mN[0][i], cur.n[1], mN[1][i], cur.n[2], mN[2][i]);
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

    /* consistent tangent */
    double* ct = ct_o + pt * 36;
    if (!flag) {
      const double AB = lam + twomu;
      for (int j = 0; j < 6; ++j)
        for (int i = 0; i < 6; ++i) ct[j * 6 + i] = (i == j) ? ((i < 3) ? AB : twomu) : ((i < 3 && j < 3) ? lam : 0.0);
    } else {
      double Cf[6], det, z[3];
      const double c = twomu * dp, iphi = cur.iphi;
      hosford_system(&cur, c, am1 * cur.iphi, Cf, &det);
      sym3_apply(Cf, cur.n, z);
      const double idet = 1.0 / det;
      for (int k = 0; k < 3; ++k) z[k] = z[k] * idet;
      const double nz = dot3(cur.n[0], z[0], cur.n[1], z[1], cur.n[2], z[2]);
      const double w = (twomu * twomu) / FMA(twomu, nz, cur.dsy); /* (2 mu z)(2 mu z)^T / (2 mu n.z + sigma_Y'(p)) */
      const double ti = twomu * idet;
      /* normal block An (symmetric): 2 mu A^-1 + lam - w z z^T */
      double An[3][3];
      An[0][0] = FNMA(w, z[0] * z[0], FMA(ti, Cf[0], lam));
      An[0][1] = FNMA(w, z[0] * z[1], FMA(ti, Cf[1], lam));
      An[0][2] = FNMA(w, z[0] * z[2], FMA(ti, Cf[2], lam));
      An[1][1] = FNMA(w, z[1] * z[1], FMA(ti, Cf[3], lam));
      An[1][2] = FNMA(w, z[1] * z[2], FMA(ti, Cf[4], lam));
      An[2][2] = FNMA(w, z[2] * z[2], FMA(ti, Cf[5], lam));
      An[1][0] = An[0][1];
      An[2][0] = An[0][2];
      An[2][1] = An[1][2];
      /* shear moduli of the pairs (0,1), (1,2), (2,0): theta = (h_k + DD/2) / phi */
      const double th01 = FMA(0.5, divdiff(-cur.u[2], cur.u[1], a), cur.h[0]) * iphi;
      const double th12 = FMA(0.5, divdiff(-cur.u[0], cur.u[2], a), cur.h[1]) * iphi;
      const double th20 = FMA(0.5, divdiff(-cur.u[1], cur.u[0], a), cur.h[2]) * iphi;
      const double G[3] = {twomu / FMA(c, th01, 1.0), twomu / FMA(c, th12, 1.0), twomu / FMA(c, th20, 1.0)};
      /* unit Mandel vectors of sym(e_i e_j), pairs in the same order */
      static const int PI[3] = {0, 1, 2}, PJ[3] = {1, 2, 0};
      double mS[3][6];
      for (int p = 0; p < 3; ++p) {
        const int i = PI[p], j = PJ[p];
        mS[p][0] = SQRT2 * (Q[0][i] * Q[0][j]);
        mS[p][1] = SQRT2 * (Q[1][i] * Q[1][j]);
        mS[p][2] = SQRT2 * (Q[2][i] * Q[2][j]);
        mS[p][3] = FMA(Q[0][i], Q[1][j], Q[1][i] * Q[0][j]);
        mS[p][4] = FMA(Q[0][i], Q[2][j], Q[2][i] * Q[0][j]);
        mS[p][5] = FMA(Q[1][i], Q[2][j], Q[2][i] * Q[1][j]);
      }
      double wN[3][6]; /* wN_i = sum_j An_ij mN_j */
      for (int i = 0; i < 3; ++i)
        for (int cc = 0; cc < 6; ++cc) wN[i][cc] = dot3(An[i][0], mN[0][cc], An[i][1], mN[1][cc], An[i][2], mN[2][cc]);
      double gS[3][6]; /* gS_k = G_k mS_k: the tangent is Q^T D Q, Q = rows (mN, mS), D = blockdiag(An, diag G) */
      for (int k = 0; k < 3; ++k)
        for (int cc = 0; cc < 6; ++cc) gS[k][cc] = G[k] * mS[k][cc];
      for (int j = 0; j < 6; ++j)
        for (int i = j; i < 6; ++i) {
          const double vs = FMA(mS[2][j], gS[2][i], FMA(mS[1][j], gS[1][i], mS[0][j] * gS[0][i]));
          const double v = FMA(mN[2][j], wN[2][i], FMA(mN[1][j], wN[1][i], FMA(mN[0][j], wN[0][i], vs)));
          ct[j * 6 + i] = v;
          ct[i * 6 + j] = v;
        }
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
