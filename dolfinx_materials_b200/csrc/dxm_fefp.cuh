// Finite-strain hot path: multiplicative (Fe.Fp) J2 plasticity, internal state be_bar (isochoric elastic
// left Cauchy-Green, Mandel 6) + p.  Replaces the per-point arithmetic behind JAXMaterial.integrate for
// jaxmat `FeFpJ2Plasticity` (dolfinx_materials/jaxmat.py:166-193, 208-234; workload
// tests/test_FeFp_jax.py:6-33, demos/jax/finite_strain_elastoplasticity/finite_strain_elastoplasticity.py:158-181).
// Operation order == oracle/fefp.py (see there for the derivation: the reference's 7-unknown local
// system reduces exactly to a 2x2 Newton in (dp, t) with be = alpha D + t 1).
//
// Layout: SoA, one Gauss point per thread.  Per point the kernel reads 25 doubles (F 9, F_old 9, p 1,
// be_bar 6) and writes 97 (PK1 9, p 1, be_bar 6, Ct 81): 976 B, coalesced streaming accesses.  The 9x9
// tangent is assembled column by column in registers from closed-form pieces and stored entry by entry.
#pragma once
#include <atomic>
#include <cstdlib>
#include <string>

#include "dxm_canon.cuh"

namespace dxm {

struct FeFpArgs {
  const double* F;  // s1 gradient buffer [9][ld]
  double* P;
  double* p;
  double* be;
  double* ct;  // [81][ld]
  const double* F_old;
  const double* p_old;
  const double* be_old;
  int64_t ld, start, count;
  bool perpoint;
  double E, mu, kappa, sig0, H, dsu, b;
  const double* pp[6];  // per-point E, nu, sig0, H, sigu, b
  StatSink stats;  // per-call statistics (dxm_canon.cuh)
  int vote;
  uint8_t* d_flag;
  int32_t* d_iter;
  double* d_resid;
  uint8_t* d_fail;
};

constexpr int kFeNewtonCap = 25;
constexpr double kFeNewtonRtol = 1e-12;
constexpr double kRsqrt2 = 0.70710678118654752440;
constexpr double kSqrt2 = 1.41421356237309504880;

// position of tensor component (i,j) in the reference's 9-vector [11,22,33,12,21,13,31,23,32]
// (dolfinx_materials/utils.py:173-186)
DXM_HD constexpr int idx9(int i, int j) {
  return i == j ? i : (i == 0 && j == 1) ? 3 : (i == 1 && j == 0) ? 4 : (i == 0 && j == 2) ? 5
                  : (i == 2 && j == 0) ? 6 : (i == 1 && j == 2) ? 7 : 8;
}

constexpr double kThird = 1.0 / 3.0;

// x^(-1/3), division free, from exactly rounded operations only (twin of oracle.fefp.rcbrt_c)
DXM_HD double rcbrt_c(double x) {
#ifdef __CUDA_ARCH__
  if (!(x > 0.0)) return __longlong_as_double(0x7ff8000000000000LL);
#else
  if (!(x > 0.0)) return NAN;
#endif
  int e;
  const double m = frexp(x, &e);
  const int q = (e >= 0) ? (e / 3) : -((-e + 2) / 3);
  const int r = e - 3 * q;
  const double xr = m * (double)(1 << r);
  double y = fnma_c(0.15, xr, 1.2);
#pragma unroll
  for (int i = 0; i < 6; ++i) y = (y * fnma_c(xr, (y * y) * y, 4.0)) * kThird;
#ifdef __CUDA_ARCH__
  return y * __hiloint2double((1023 - q) << 20, 0);
#else
  return ldexp(y, -q);  // exact: y in [0.7, 1.3], |q| <= 358
#endif
}

DXM_HD double dot3(double a0, double b0, double a1, double b1, double a2,
                                       double b2) {
  return fma_c(a2, b2, fma_c(a1, b1, a0 * b0));
}

DXM_HD double det3(const double (&A)[3][3]) {
  const double m0 = fms_c(A[1][1], A[2][2], A[1][2] * A[2][1]);
  const double m1 = fms_c(A[1][0], A[2][2], A[1][2] * A[2][0]);
  const double m2 = fms_c(A[1][0], A[2][1], A[1][1] * A[2][0]);
  return fma_c(A[0][2], m2, fnma_c(A[0][1], m1, A[0][0] * m0));
}

DXM_HD void inv3(const double (&A)[3][3], double (&Ai)[3][3], double& det) {
  double c[3][3];
  c[0][0] = fms_c(A[1][1], A[2][2], A[1][2] * A[2][1]);
  c[0][1] = fms_c(A[0][2], A[2][1], A[0][1] * A[2][2]);
  c[0][2] = fms_c(A[0][1], A[1][2], A[0][2] * A[1][1]);
  c[1][0] = fms_c(A[1][2], A[2][0], A[1][0] * A[2][2]);
  c[1][1] = fms_c(A[0][0], A[2][2], A[0][2] * A[2][0]);
  c[1][2] = fms_c(A[0][2], A[1][0], A[0][0] * A[1][2]);
  c[2][0] = fms_c(A[1][0], A[2][1], A[1][1] * A[2][0]);
  c[2][1] = fms_c(A[0][1], A[2][0], A[0][0] * A[2][1]);
  c[2][2] = fms_c(A[0][0], A[1][1], A[0][1] * A[1][0]);
  det = fma_c(A[0][2], c[2][0], fma_c(A[0][1], c[1][0], A[0][0] * c[0][0]));
  const double rdet = 1.0 / det;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Ai[i][j] = c[i][j] * rdet;
}

struct FeFpLocalProps {
  double threemu, sig0, H, dsu, bdsu, b;
};

// local 2x2 Newton in (dp, t); lanes outside `mask` must not call.  Inputs are six scalars per point, which is
// what makes the block-level compaction below cheap.
DXM_HD void fefp_newton(const FeFpLocalProps& m, const double seq, const double dd,
                                            const double d3, const double p_old, double& ecur, double& dp,
                                            double& t, int& n_iter, double& resid, bool& fail, bool active,
                                            const unsigned mask, const bool vote) {
  const double c = m.threemu * (1.0 / seq);
  const double tol1 = kFeNewtonRtol * seq;
  for (int it = 0; vote ? DXM_ANY_SYNC(mask, active) : active; ++it) {
    if (active) {
      const double ct = c * t;
      const double tmt = m.threemu * t;
      const double alpha = fnma_c(ct, dp, 1.0);
      const double p = p_old + dp;
      const double sy = fma_c(m.dsu, 1.0 - ecur, fma_c(m.H, p, m.sig0));
      const double r1 = fnma_c(tmt, dp, seq) - sy;
      const double a2 = alpha * alpha;
      const double ha2 = 0.5 * a2;
      const double tt = t * t;
      const double r2 = fnma_c(ha2, dd * t, tt * t) + fms_c(a2 * alpha, d3, 1.0);
      if (fabs(r1) <= tol1 && fabs(r2) <= kFeNewtonRtol) {
        resid = fabs(r1);
        active = false;
      } else if (it == kFeNewtonCap) {
        resid = fabs(r1);
        fail = true;
        active = false;
      } else {
        const double dsy = fma_c(m.bdsu, ecur, m.H);
        const double g = fnma_c(alpha * dd, t, 3.0 * (a2 * d3));
        const double J11 = -tmt - dsy;
        const double J12 = -(m.threemu * dp);
        const double cdp = c * dp;
        const double J21 = -(g * ct);
        const double J22 = fnma_c(g, cdp, fnma_c(ha2, dd, 3.0 * tt));
        const double rdet = 1.0 / fms_c(J11, J22, J12 * J21);
        const double dp_new = fma_c(fms_c(J12, r2, r1 * J22), rdet, dp);
        const double t_new = fma_c(fms_c(J21, r1, J11 * r2), rdet, t);
        dp = dp_new;
        t = t_new;
        ecur = exp_hd(-(m.b * (p_old + dp)));
        ++n_iter;
      }
    }
  }
}

// block-level compaction of the plastic points' local solves (see CompactSmem in dxm_small_strain.cuh)
struct FeFpCompactSmem {
  double a[6][128];  // in: seq, dd, d3, t0, p_old, ecur   out: dp, t, ecur, resid
  int meta[128];
  int warp_count[4];
};
template <bool C>
struct FeFpCompactStore {
  FeFpCompactSmem s;
  __device__ __forceinline__ FeFpCompactSmem* get() { return &s; }
};
template <>
struct FeFpCompactStore<false> {
  __device__ __forceinline__ FeFpCompactSmem* get() { return nullptr; }
};

// One Gauss point at SoA position i0: loads, local solve, state / stress / tangent stores (the tangent entries are
// stored as they are formed, so the 81 of them never sit in registers together).  __host__ __device__ (without
// COMPACT) so that a CPU test can run the very code the kernel runs per point against the oracle
// (tests/point_host_check.cu) -- the product only ever calls it from the kernel below.
template <bool PERPOINT, bool DIAG, bool COMPACT>
DXM_HD void fefp_point(const FeFpArgs& a, const int64_t i0, const bool live, const unsigned warp_mask,
                       FeFpCompactSmem* cs, PointStats& acc) {
  const int64_t ld = a.ld;
  double A[3][3], Ao[3][3], Bo[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      A[i][j] = ld_stream(a.F + (int64_t)idx9(i, j) * ld + i0);
      Ao[i][j] = ld_stream(a.F_old + (int64_t)idx9(i, j) * ld + i0);
    }
  {
    double beo[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) beo[c] = ld_stream(a.be_old + (int64_t)c * ld + i0);
    Bo[0][0] = beo[0];
    Bo[1][1] = beo[1];
    Bo[2][2] = beo[2];
    Bo[0][1] = Bo[1][0] = beo[3] * kRsqrt2;
    Bo[0][2] = Bo[2][0] = beo[4] * kRsqrt2;
    Bo[1][2] = Bo[2][1] = beo[5] * kRsqrt2;
  }
  const double p_old = ld_stream(a.p_old + i0);
  // a singular or inverted elastic state (det be_bar <= 0; e.g. an all-zero be_bar that was never initialised to the
  // identity) would give PK1 = 0 with every check green: such a point is counted as failed
  const double det_bo = det3(Bo);

  double mu, kappa, sig0, H, dsu, b;
  if (PERPOINT) {
    const double E = ld_stream(a.pp[0] + i0), nu = ld_stream(a.pp[1] + i0);
    mu = E / 2.0 / (1.0 + nu);
    kappa = E / (3.0 * (1.0 - 2.0 * nu));
    sig0 = ld_stream(a.pp[2] + i0);
    H = ld_stream(a.pp[3] + i0);
    const double d = ld_stream(a.pp[4] + i0) - sig0;
    dsu = isfinite(d) ? d : 0.0;
    b = ld_stream(a.pp[5] + i0);
  } else {
    mu = a.mu;
    kappa = a.kappa;
    sig0 = a.sig0;
    H = a.H;
    dsu = a.dsu;
    b = a.b;
  }
  const double threemu = 3.0 * mu;
  const double bdsu = b * dsu;

  // ---- trial state ----------------------------------------------------------------------------
  double B[3][3], D[3][3];
  {
    double Aoi[3][3], deto, f[3][3], M[3][3];
    inv3(Ao, Aoi, deto);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        f[i][j] = dot3(A[i][0], Aoi[0][j], A[i][1], Aoi[1][j], A[i][2], Aoi[2][j]);
    const double Jf = det3(f);
    const double rc = rcbrt_c(Jf);
    const double s23 = rc * rc;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        M[i][j] = dot3(f[i][0], Bo[0][j], f[i][1], Bo[1][j], f[i][2], Bo[2][j]);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = i; j < 3; ++j) {
        B[i][j] = s23 * dot3(M[i][0], f[j][0], M[i][1], f[j][1], M[i][2], f[j][2]);
        B[j][i] = B[i][j];
      }
  }
  const double t0 = ((B[0][0] + B[1][1]) + B[2][2]) * kThird;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) D[i][j] = (i == j) ? (B[i][j] - t0) : B[i][j];
  const double dd = fma_c(2.0, dot3(D[0][1], D[0][1], D[0][2], D[0][2], D[1][2], D[1][2]),
                          dot3(D[0][0], D[0][0], D[1][1], D[1][1], D[2][2], D[2][2]));
  const double d3 = det3(D);
  const double seq = mu * sqrt(1.5 * dd);
  const double rseq = 1.0 / seq;

  double ecur = exp_hd(-(b * p_old));
  const double sy0 = fma_c(dsu, 1.0 - ecur, fma_c(H, p_old, sig0));
  const double ftr = seq - sy0;
  const bool flag = ftr > 0.0;

  // ---- local 2x2 Newton in (dp, t) ----------------------------------------------------------------
  const double c = threemu * rseq;
  double dp = 0.0, t = t0, resid = 0.0;
  int n_iter = 0;
  bool fail = false;
  {
    FeFpLocalProps lp;
    lp.threemu = threemu;
    lp.sig0 = sig0;
    lp.H = H;
    lp.dsu = dsu;
    lp.bdsu = bdsu;
    lp.b = b;
    const bool active = live && flag;
    if (!COMPACT) {
      // warp-synchronous with a warp-vote early exit (see dxm_small_strain.cuh)
      fefp_newton(lp, seq, dd, d3, p_old, ecur, dp, t, n_iter, resid, fail, active, warp_mask, a.vote != 0);
    } else {
#ifdef __CUDA_ARCH__
      const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
      const unsigned bal = __ballot_sync(0xffffffffu, active);
      if (lane == 0) cs->warp_count[w] = __popc(bal);
      __syncthreads();
      int base = 0, total = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int cnt = cs->warp_count[i];
        base += i < w ? cnt : 0;
        total += cnt;
      }
      const int slot = base + __popc(bal & ((1u << lane) - 1u));
      if (active) {
        cs->a[0][slot] = seq;
        cs->a[1][slot] = dd;
        cs->a[2][slot] = d3;
        cs->a[3][slot] = t0;
        cs->a[4][slot] = p_old;
        cs->a[5][slot] = ecur;
      }
      __syncthreads();
      const bool solver = (int)threadIdx.x < total;
      const unsigned smask = __ballot_sync(0xffffffffu, solver);
      if (solver) {
        const int k = threadIdx.x;
        const double sq = cs->a[0][k], dd_s = cs->a[1][k], d3_s = cs->a[2][k], po = cs->a[4][k];
        double ts = cs->a[3][k], ec = cs->a[5][k], d = 0.0, rs = 0.0;
        int ni = 0;
        bool fl = false;
        fefp_newton(lp, sq, dd_s, d3_s, po, ec, d, ts, ni, rs, fl, true, smask, a.vote != 0);
        cs->a[0][k] = d;
        cs->a[1][k] = ts;
        cs->a[2][k] = ec;
        cs->a[3][k] = rs;
        cs->meta[k] = ni | (fl ? 1 << 16 : 0);
      }
      __syncthreads();
      if (active) {
        dp = cs->a[0][slot];
        t = cs->a[1][slot];
        ecur = cs->a[2][slot];
        resid = cs->a[3][slot];
        const int mt = cs->meta[slot];
        n_iter = mt & 0xffff;
        fail = (mt >> 16) != 0;
      }
      __syncthreads();  // slots are reused by the next tile
      if (!live) return;
#endif
    }
  }
  const double alpha = flag ? fnma_c(c * t, dp, 1.0) : 1.0;
  const double p_new = p_old + dp;

  // ---- new state -----------------------------------------------------------------------------------
  double be[6];
#pragma unroll
  for (int i = 0; i < 3; ++i) be[i] = flag ? fma_c(alpha, D[i][i], t) : B[i][i];
  be[3] = (alpha * D[0][1]) * kSqrt2;
  be[4] = (alpha * D[0][2]) * kSqrt2;
  be[5] = (alpha * D[1][2]) * kSqrt2;

  // ---- stress: tau = mu alpha D + pvol 1,  PK1 = tau F^-T = mu alpha (D F^-T) + pvol F^-T --------------
  double Ai[3][3], Jd, DA[3][3], P[3][3];
  inv3(A, Ai, Jd);
  const double muA = mu * alpha;
  const double pvol = (0.5 * kappa) * fms_c(Jd, Jd, 1.0);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      DA[i][j] = dot3(D[i][0], Ai[j][0], D[i][1], Ai[j][1], D[i][2], Ai[j][2]);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) P[i][j] = fma_c(muA, DA[i][j], pvol * Ai[j][i]);

  double chk = (seq + fabs(Jd)) + p_new;
#pragma unroll
  for (int i = 0; i < 6; ++i) chk = chk + fabs(be[i]);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) chk = chk + fabs(P[i][j]);
  if (!isfinite(chk) || !(det_bo > 0.0)) fail = true;

  // state and stress stores (the tangent follows)
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) st_stream(a.P + (int64_t)idx9(i, j) * ld + i0, P[i][j]);
  st_stream(a.p + i0, p_new);
#pragma unroll
  for (int c6 = 0; c6 < 6; ++c6) st_stream(a.be + (int64_t)c6 * ld + i0, be[c6]);

  acc.n_plastic += flag ? 1u : 0u;
  acc.n_fail += fail ? 1u : 0u;
  acc.max_iter = n_iter > (int)acc.max_iter ? (unsigned)n_iter : acc.max_iter;
  acc.max_resid = resid > acc.max_resid ? resid : acc.max_resid;
  if (resid != resid) acc.max_resid = resid;
  if (DIAG) {
    a.d_flag[i0] = flag ? 1 : 0;
    a.d_iter[i0] = n_iter;
    a.d_resid[i0] = resid;
    a.d_fail[i0] = fail ? 1 : 0;
  }

  // ---- local-solve sensitivities: d(alpha) = al1 (D:dD) + al2 (D^2:dD) --------------------------------
  double al1 = 0.0, al2 = 0.0;
  if (flag) {
    const double sq1 = (1.5 * (mu * mu)) * rseq;
    const double a2 = alpha * alpha;
    const double dsy = fma_c(bdsu, ecur, H);
    const double g = fnma_c(alpha * dd, t, 3.0 * (a2 * d3));
    const double ct = c * t;
    const double cdp = c * dp;
    const double J11 = -(threemu * t) - dsy;
    const double J12 = -(threemu * dp);
    const double J21 = -(g * ct);
    const double J22 = fnma_c(g, cdp, fnma_c(0.5 * a2, dd, 3.0 * (t * t)));
    const double rdet = 1.0 / fms_c(J11, J22, J12 * J21);
    const double oma = (1.0 - alpha) * rseq;
    const double b21 = fms_c(g * oma, sq1, a2 * t);
    const double b22 = a2 * alpha;
    const double p1 = -(fms_c(sq1, J22, J12 * b21) * rdet);
    const double t1 = -(fms_c(J11, b21, J21 * sq1) * rdet);
    const double p2 = (J12 * b22) * rdet;
    const double t2 = -((J11 * b22) * rdet);
    al1 = fnma_c(cdp, t1, fnma_c(ct, p1, oma * sq1));
    al2 = fnma_c(cdp, t2, -(ct * p2));
  }

  // ---- tangent, column (k,l) = d/dF_kl, row (i,j) = PK1_ij; w = row l of F^-1, v = B w = D w + t0 w:
  //      dP_ij = cD (D F^-T)_ij + cI F^-T_ij + delta_ik mu alpha (F^-1 v)_j + hs w_i F^-1_jk
  const double kJ2 = kappa * (Jd * Jd);
  const double c23dd = (2.0 / 3.0) * dd;
  const double twod3 = 2.0 * d3;
  const double c23muA = (2.0 / 3.0) * muA;
  const double hs = fms_c(muA, t0, pvol);
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    double w[3], v[3], u[3], z[3], my[3], hw[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) w[i] = Ai[l][i];
#pragma unroll
    for (int i = 0; i < 3; ++i) v[i] = fma_c(t0, w[i], dot3(D[i][0], w[0], D[i][1], w[1], D[i][2], w[2]));
    if (flag) {
#pragma unroll
      for (int i = 0; i < 3; ++i) u[i] = dot3(D[i][0], v[0], D[i][1], v[1], D[i][2], v[2]);
#pragma unroll
      for (int i = 0; i < 3; ++i) z[i] = dot3(D[i][0], u[0], D[i][1], u[1], D[i][2], u[2]);
    } else {
#pragma unroll
      for (int i = 0; i < 3; ++i) u[i] = z[i] = 0.0;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) my[j] = muA * dot3(Ai[j][0], v[0], Ai[j][1], v[1], Ai[j][2], v[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) hw[i] = hs * w[i];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double cd0 = 0.0;
      if (flag) {
        const double a1 = fnma_c(c23dd, w[k], 2.0 * u[k]);
        const double a2p = fnma_c(twod3, w[k], fnma_c(c23dd, v[k], 2.0 * z[k]));
        cd0 = mu * fma_c(al2, a2p, al1 * a1);
      }
      const double cD = fnma_c(c23muA, w[k], cd0);
      const double cI = fnma_c(c23muA, v[k], kJ2 * w[k]);
      const int col = idx9(k, l);
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          double val = fma_c(hw[i], Ai[j][k], fma_c(cI, Ai[j][i], cD * DA[i][j]));
          if (i == k) val = val + my[j];
          st_stream(a.ct + (int64_t)(idx9(i, j) * 9 + col) * ld + i0, val);
        }
    }
  }
}

template <bool PERPOINT, bool DIAG, int MINB, bool COMPACT = false>
__global__ void __launch_bounds__(128, MINB)
    dxm_fefp_kernel(const FeFpArgs a) {
  static_assert(!COMPACT || !PERPOINT, "compaction: uniform properties only");
  __shared__ FeFpCompactStore<COMPACT> cs_storage;
  FeFpCompactSmem* cs = cs_storage.get();
  const int64_t ntile = (a.count + blockDim.x - 1) / blockDim.x;
  PointStats acc;

  for (int64_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int64_t loc = tile * blockDim.x + threadIdx.x;
    const unsigned warp_mask = __ballot_sync(0xffffffffu, loc < a.count);
    const bool live = loc < a.count;
    if (!COMPACT && !live) continue;
    // COMPACT: every thread of the CTA takes part in the block-level exchange; dead lanes shadow point 0
    const int64_t i0 = a.start + (live ? loc : 0);
    fefp_point<PERPOINT, DIAG, COMPACT>(a, i0, live, warp_mask, cs, acc);
  }
  block_reduce_stats(acc, a.stats);
}

template <int MINB>
inline const void* fefp_kernel_ptr(bool perpoint, bool diag, bool compact = false) {
  if (compact && !perpoint && MINB == 4)
    return diag ? (const void*)dxm_fefp_kernel<false, true, 4, true> : (const void*)dxm_fefp_kernel<false, false, 4, true>;
  return perpoint ? (diag ? (const void*)dxm_fefp_kernel<true, true, MINB>
                          : (const void*)dxm_fefp_kernel<true, false, MINB>)
                  : (diag ? (const void*)dxm_fefp_kernel<false, true, MINB>
                          : (const void*)dxm_fefp_kernel<false, false, MINB>);
}

inline int launch_fefp(const FeFpArgs& a, bool diag, bool compact, int num_sms, cudaStream_t stream,
                       std::atomic<long long>* launches, std::string* err) {
  // small batches: 64-point CTAs of one tile each over all SMs (latency-bound; see launch_small_strain)
  const bool small = !compact && a.count <= (int64_t)num_sms * 128;
  const int block = small ? 64 : 128;
  const int64_t ntile = (a.count + block - 1) / block;
  // DXM_FEFP_MINB: resident CTAs per SM the register allocation targets (3: <=168 regs, 4: <=128, 5: <=96)
  static const int minb = [] {
    const char* e = std::getenv("DXM_FEFP_MINB");
    return e ? std::atoi(e) : 4;
  }();
  const void* k = minb == 4 ? fefp_kernel_ptr<4>(a.perpoint, diag, compact)
                  : minb == 5 ? fefp_kernel_ptr<5>(a.perpoint, diag)
                              : fefp_kernel_ptr<3>(a.perpoint, diag);
  // a few tiles per CTA (see grid_for in dxm_api.cu); DXM_GRID=<k> forces k CTAs per SM
  static const int mult = [] {
    const char* e = std::getenv("DXM_GRID");
    return e ? std::atoi(e) : 0;
  }();
  static const int tpb = [] {
    const char* e = std::getenv("DXM_TPB");
    const int v = e ? std::atoi(e) : 2;
    return v > 0 ? v : 2;
  }();
  int64_t grid = small ? ntile : mult > 0 ? (int64_t)mult * num_sms : (ntile + tpb - 1) / tpb;
  if (grid > ntile) grid = ntile;
  if (grid > 0x7fffffff) grid = 0x7fffffff;
  if (grid < 1) grid = 1;
  void* args[] = {(void*)&a};
  cudaError_t e = cudaLaunchKernel(k, dim3((unsigned)grid), dim3(block), args, 0, stream);
  launches->fetch_add(1);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    *err = std::string("dxm_fefp_kernel launch: ") + cudaGetErrorString(e);
    return -1;
  }
  return 0;
}

}  // namespace dxm
