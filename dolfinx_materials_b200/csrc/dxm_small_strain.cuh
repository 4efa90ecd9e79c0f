// Small-strain hot path: linear elasticity and J2 plasticity (linear / Voce / mixed hardening).
//
// Replaces the per-point arithmetic behind JAXMaterial.integrate (dolfinx_materials/jaxmat.py:208-234,
// jaxmat `vonMisesIsotropicHardening`) and LinearElasticIsotropic.constitutive_update
// (python_materials/elasticity.py:21-24); closed form per
// tests/mfront/IsotropicLinearHardeningPlasticity.mfront:49-77.  Operation order == oracle/small_strain.py.
//
// Layout: SoA, one Gauss point per thread-lane, PPT consecutive points per thread (PPT=2 -> 16-byte
// double2 accesses).  Per point the kernel reads 25 doubles (eps 6, eps_old 6, sig_old 6, p_old 1,
// epsp_old 6) and writes 34 (sig 6, p 1, epsp 6, and the 21 unique entries of the symmetric Ct): 472 B of
// traffic for 592 algorithmic bytes (the full 36-entry tangent of the reference boundary), all coalesced streaming.
// The tangent is never held in registers: Ct = A 1x1 + B I - gamma n x n is formed entry by entry at
// store time from 3 scalars and the flow direction.
#pragma once
#include "dxm_canon.cuh"

namespace dxm {

constexpr int kNewtonCap = 25;
constexpr double kNewtonRtol = 1e-12;

enum Hardening { HARD_NONE = 0, HARD_LINEAR = 1, HARD_GENERAL = 2, HARD_TABLE = 3 };
constexpr int kMaxTable = 64;  // points of a piecewise-linear hardening table

struct SmallStrainArgs {
  // s1 (written) -- eps is read from s1.strain (the gradient buffer)
  const double* eps;
  double* sig;
  double* p;
  double* epsp;
  double* ct;
  // s0 (read)
  const double* eps_old;
  const double* sig_old;
  const double* p_old;
  const double* epsp_old;
  int64_t ld;     // SoA leading dimension
  int64_t start;  // first point of this launch (multiple of 2)
  int64_t count;  // number of points
  // uniform properties (PERPOINT == false)
  double lam, mu, sig0, H, dsu, b;
  // per-point properties (PERPOINT == true), each [ld]
  const double *pE, *pnu, *psig0, *pH, *psigu, *pb;
  // HARD_TABLE: piecewise-linear sigma_Y through (tp[k], ts[k]) with segment slopes tH[k] (last repeated), [3][ntab]
  const double* table;
  int ntab;
  int hos_a;  // DXM_HOSFORD_LINEAR: exponent of the Hosford criterion (even integer)
  double hos_bound;  // ... and sup sigma_eq / seq_Mises (1 + 1e-9): points below it are finished without a local solve
  StatSink stats;  // per-call statistics (dxm_canon.cuh)
  int vote;  // 1: warp-vote (__any_sync) Newton loop exit, 0: per-lane exit (A/B knob DXM_VOTE)
  // optional diagnostics (DIAG == true)
  uint8_t* d_flag;
  int32_t* d_iter;
  double* d_resid;
  uint8_t* d_fail;
};

struct PointProps {
  double lam, mu, sig0, H, dsu, b;
  const double *tp, *ts, *tH;  // HARD_TABLE (shared memory)
  int ntab;
};

// ---- block-level compaction of the local Newton solves ("warp compaction of plastic points") -------------------
// The Voce Newton depends on three scalars per point (seq, p_old, exp(-b p_old)).  With COMPACT the plastic
// points of a 256-point tile publish them to consecutive shared-memory slots (ballot + popc prefix inside the
// warp, warp totals across the CTA), threads 0..n_plastic-1 each solve one slot, and the owners read dp back:
// only ceil(n_plastic/32) warps execute the loop instead of every warp that holds at least one plastic lane.
// Same arithmetic per point -> bit-identical results.  Uniform properties only.
struct CompactSmem {
  double a[3][256];      // in: seq, p_old, ecur   out: dp, ecur, resid
  int meta[256];         // out: n_iter | fail << 16
  int warp_count[8];
};

template <bool C>
struct CompactStore {
  CompactSmem s;
  __device__ __forceinline__ CompactSmem* get() { return &s; }
};
template <>
struct CompactStore<false> {
  __device__ __forceinline__ CompactSmem* get() { return nullptr; }
};

// capped scalar Newton on r(dp) = seq - 3 mu dp - sigY(p_old + dp); lanes outside `mask` must not call
DXM_HD void voce_newton(const PointProps& m, const double threemu, const double bdsu,
                                            const double seq, const double p_old, double& ecur, double& dp,
                                            int& n_iter, double& resid, bool& fail, bool active,
                                            const unsigned mask, const bool vote) {
  const double tol = kNewtonRtol * seq;
  for (int it = 0; vote ? DXM_ANY_SYNC(mask, active) : active; ++it) {
    if (active) {
      const double p = p_old + dp;
      const double sy = fma_c(m.dsu, 1.0 - ecur, fma_c(m.H, p, m.sig0));
      const double r = fnma_c(threemu, dp, seq) - sy;
      if (fabs(r) <= tol) {
        resid = fabs(r);
        active = false;
      } else if (it == kNewtonCap) {
        resid = fabs(r);
        fail = true;
        active = false;
      } else {
        const double dsy = fma_c(bdsu, ecur, m.H);
        dp = dp + r / (threemu + dsy);
        ecur = exp_hd(-(m.b * (p_old + dp)));
        ++n_iter;
      }
    }
  }
}

// One Gauss point.  Returns results through references; everything stays in registers.  __host__ __device__ (without
// COMPACT) so that a CPU test can run the very code the kernel runs per point against the oracle
// (tests/point_host_check.cu) -- the product only ever calls it from the kernel below.
template <int HARD, bool COMPACT>
DXM_HD void j2_point(const PointProps& m, const double (&eps)[6],
                                         const double (&e_old)[6], const double (&s_old)[6],
                                         const double p_old, const double (&ep_old)[6],
                                         double (&sig)[6], double& p_new, double (&epsp)[6],
                                         double (&nrm)[6], double& A, double& B, double& gamma,
                                         bool& flag, int& n_iter, double& resid, bool& fail,
                                         const unsigned warp_mask, const bool vote, const bool live,
                                         CompactSmem* cs) {
  const double twomu = 2.0 * m.mu;
  const double threemu = 3.0 * m.mu;
  double de[6], st[6], s[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) de[i] = eps[i] - e_old[i];
  const double tr = (de[0] + de[1]) + de[2];
  const double ltr = m.lam * tr;
#pragma unroll
  for (int i = 0; i < 3; ++i) st[i] = s_old[i] + fma_c(twomu, de[i], ltr);
#pragma unroll
  for (int i = 3; i < 6; ++i) st[i] = fma_c(twomu, de[i], s_old[i]);
  const double pm = ((st[0] + st[1]) + st[2]) / 3.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) s[i] = st[i] - pm;
#pragma unroll
  for (int i = 3; i < 6; ++i) s[i] = st[i];
  double ss = s[0] * s[0];
#pragma unroll
  for (int i = 1; i < 6; ++i) ss = fma_c(s[i], s[i], ss);
  const double seq = sqrt(1.5 * ss);

  double dp = 0.0;
  double Hp = m.H;
  flag = false;
  n_iter = 0;
  resid = 0.0;
  fail = false;

  if (HARD != HARD_NONE) {
    const double bdsu = m.b * m.dsu;
    double ecur = 1.0;
    double sy0;
    int seg = 0;
    if (HARD == HARD_TABLE) {
      for (int k = 0; k + 1 < m.ntab; ++k)
        if (p_old >= m.tp[k + 1]) seg = k + 1;
      sy0 = fma_c(m.tH[seg], p_old - m.tp[seg], m.ts[seg]);
    } else if (HARD == HARD_GENERAL) {
      ecur = exp_hd(-(m.b * p_old));
      sy0 = fma_c(m.dsu, 1.0 - ecur, fma_c(m.H, p_old, m.sig0));
    } else {
      sy0 = fma_c(m.H, p_old, m.sig0);
    }
    const double f = seq - sy0;
    flag = f > 0.0;
    // closed-form radial return (mfront :57-60) for linear hardening / no saturation term
    const bool closed = (HARD == HARD_LINEAR) || (HARD == HARD_TABLE) || (bdsu == 0.0);
    if (HARD == HARD_TABLE) {
      // exact walk over the segments of the piecewise-linear curve: closed form per segment, move on while the
      // solution leaves the segment (n_iter counts crossings)
      if (live && flag) {
        for (;;) {
          dp = (seq - fma_c(m.tH[seg], p_old - m.tp[seg], m.ts[seg])) / (threemu + m.tH[seg]);
          if (seg + 1 < m.ntab && p_old + dp > m.tp[seg + 1]) {
            ++seg;
            ++n_iter;
          } else {
            break;
          }
        }
      }
      Hp = m.tH[seg];
    } else if (flag && closed) {
      dp = f / (threemu + m.H);
    }
    if (HARD == HARD_GENERAL) {
      // capped scalar Newton, warp-synchronous: every lane of the warp stays in the loop until the
      // warp vote says no lane is still iterating (early exit as soon as the slowest lane converged)
      const bool active = live && flag && !closed;
      if (!COMPACT) {
        voce_newton(m, threemu, bdsu, seq, p_old, ecur, dp, n_iter, resid, fail, active, warp_mask, vote);
      } else {
#ifdef __CUDA_ARCH__
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        const unsigned bal = __ballot_sync(0xffffffffu, active);
        if (lane == 0) cs->warp_count[w] = __popc(bal);
        __syncthreads();
        int base = 0, total = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = cs->warp_count[i];
          base += i < w ? c : 0;
          total += c;
        }
        const int slot = base + __popc(bal & ((1u << lane) - 1u));
        if (active) {
          cs->a[0][slot] = seq;
          cs->a[1][slot] = p_old;
          cs->a[2][slot] = ecur;
        }
        __syncthreads();
        const bool solver = (int)threadIdx.x < total;
        const unsigned smask = __ballot_sync(0xffffffffu, solver);
        if (solver) {
          const double sq = cs->a[0][threadIdx.x], po = cs->a[1][threadIdx.x];
          double ec = cs->a[2][threadIdx.x], d = 0.0, rs = 0.0;
          int ni = 0;
          bool fl = false;
          voce_newton(m, threemu, bdsu, sq, po, ec, d, ni, rs, fl, true, smask, vote);
          cs->a[0][threadIdx.x] = d;
          cs->a[2][threadIdx.x] = ec;
          cs->a[1][threadIdx.x] = rs;
          cs->meta[threadIdx.x] = ni | (fl ? 1 << 16 : 0);
        }
        __syncthreads();
        if (active) {
          dp = cs->a[0][slot];
          ecur = cs->a[2][slot];
          resid = cs->a[1][slot];
          const int mt = cs->meta[slot];
          n_iter = mt & 0xffff;
          fail = (mt >> 16) != 0;
        }
        __syncthreads();  // slots are reused by the next tile
#endif
      }
    }
    if (HARD == HARD_GENERAL) Hp = fma_c(bdsu, ecur, m.H);
  }

  double q = 0.0;
  gamma = 0.0;
  if (flag) {
#pragma unroll
    for (int i = 0; i < 6; ++i) nrm[i] = (1.5 * s[i]) / seq;
    q = dp / seq;
  } else {
#pragma unroll
    for (int i = 0; i < 6; ++i) nrm[i] = 0.0;
    dp = 0.0;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const double depsp = dp * nrm[i];
    sig[i] = fnma_c(twomu, depsp, st[i]);
    epsp[i] = ep_old[i] + depsp;
  }
  p_new = p_old + dp;

  const double fourmu2 = (4.0 * m.mu) * m.mu;
  const double beta = fourmu2 * q;
  if (flag) {
    const double cste = 1.0 / (threemu + Hp);
    gamma = fourmu2 * (cste - q);
  }
  A = fma_c(0.5, beta, m.lam);
  B = fnma_c(1.5, beta, twomu);

  double chk = (seq + fabs(pm)) + p_new;
#pragma unroll
  for (int i = 0; i < 6; ++i) chk = chk + fabs(epsp[i]);
  if (!isfinite(chk)) fail = true;
}

// Lame constants and hardening parameters of one point from its (E, nu, sig0, H, sigu, b) row (per-point properties)
DXM_HD void point_props(const double E, const double nu, const double sig0, const double H, const double sigu,
                        const double b, PointProps& m) {
  m.lam = E * nu / (1.0 + nu) / (1.0 - 2.0 * nu);
  m.mu = E / 2.0 / (1.0 + nu);
  m.sig0 = sig0;
  m.H = H;
  const double d = sigu - sig0;
  m.dsu = isfinite(d) ? d : 0.0;
  m.b = b;
}

// entry (j, i), j <= i, of the symmetric tangent Ct = A 1x1 + B I - gamma n x n
DXM_HD double j2_tangent_entry(const int i, const int j, const double A, const double B, const double gamma,
                               const double ni, const double nj) {
  double base;
  if (i == j)
    base = (i < 3) ? (A + B) : B;
  else if (i < 3 && j < 3)
    base = A;
  else
    base = 0.0;
  return fnma_c(gamma, ni * nj, base);
}

template <int HARD, bool PERPOINT, int PPT, bool DIAG, int MINB, bool COMPACT = false>
__global__ void __launch_bounds__(256, MINB)
    dxm_small_strain_kernel(const SmallStrainArgs a) {
  static_assert(!COMPACT || (HARD == HARD_GENERAL && !PERPOINT && PPT == 1), "compaction: uniform Voce, PPT = 1");
  static_assert(HARD != HARD_TABLE || PPT == 1, "tabulated hardening: PPT = 1");
  __shared__ CompactStore<COMPACT> cs_storage;
  CompactSmem* cs = cs_storage.get();
  __shared__ double s_table[HARD == HARD_TABLE ? 3 * kMaxTable : 1];
  if (HARD == HARD_TABLE) {
    for (int i = threadIdx.x; i < 3 * a.ntab; i += blockDim.x) s_table[i] = a.table[i];
    __syncthreads();
  }
  const int64_t ld = a.ld;
  const int64_t ntile = (a.count + (int64_t)blockDim.x * PPT - 1) / ((int64_t)blockDim.x * PPT);
  PointStats acc;

  for (int64_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int64_t loc = (tile * blockDim.x + threadIdx.x) * PPT;  // local index within launch
    // lanes of this warp that own points in this tile (the vote mask of the local Newton loop)
    const unsigned warp_mask = __ballot_sync(0xffffffffu, loc < a.count);
    const bool live = loc < a.count;
    if (!COMPACT && !live) continue;
    // COMPACT: every thread of the CTA takes part in the block-level exchange; dead lanes shadow point 0
    const int64_t i0 = a.start + (live ? loc : 0);

    double eps[6][PPT], e_old[6][PPT], s_old[6][PPT], ep_old[6][PPT], p_old[PPT];
#pragma unroll
    for (int c = 0; c < 6; ++c) ldv<PPT>(a.eps + c * ld + i0, eps[c]);
#pragma unroll
    for (int c = 0; c < 6; ++c) ldv<PPT>(a.eps_old + c * ld + i0, e_old[c]);
#pragma unroll
    for (int c = 0; c < 6; ++c) ldv<PPT>(a.sig_old + c * ld + i0, s_old[c]);
    ldv<PPT>(a.p_old + i0, p_old);
#pragma unroll
    for (int c = 0; c < 6; ++c) ldv<PPT>(a.epsp_old + c * ld + i0, ep_old[c]);

    double vE[PPT], vnu[PPT], vs0[PPT], vH[PPT], vsu[PPT], vb[PPT];
    if (PERPOINT) {
      ldv<PPT>(a.pE + i0, vE);
      ldv<PPT>(a.pnu + i0, vnu);
      ldv<PPT>(a.psig0 + i0, vs0);
      ldv<PPT>(a.pH + i0, vH);
      ldv<PPT>(a.psigu + i0, vsu);
      ldv<PPT>(a.pb + i0, vb);
    }

    double sig[6][PPT], epsp[6][PPT], nrm[6][PPT], p_new[PPT], A[PPT], B[PPT], gamma[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      PointProps m;
      if (PERPOINT) {
        point_props(vE[k], vnu[k], vs0[k], vH[k], vsu[k], vb[k], m);
      } else {
        m.lam = a.lam;
        m.mu = a.mu;
        m.sig0 = a.sig0;
        m.H = a.H;
        m.dsu = a.dsu;
        m.b = a.b;
      }
      m.tp = s_table;
      m.ts = s_table + a.ntab;
      m.tH = s_table + 2 * a.ntab;
      m.ntab = a.ntab;
      double e1[6], e0[6], s0[6], ep0[6], so[6], epo[6], nn[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        e1[c] = eps[c][k];
        e0[c] = e_old[c][k];
        s0[c] = s_old[c][k];
        ep0[c] = ep_old[c][k];
      }
      bool flag, fail;
      int n_iter;
      double resid;
      j2_point<HARD, COMPACT>(m, e1, e0, s0, p_old[k], ep0, so, p_new[k], epo, nn, A[k], B[k], gamma[k],
                              flag, n_iter, resid, fail, warp_mask, a.vote != 0, live, cs);
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        sig[c][k] = so[c];
        epsp[c][k] = epo[c];
        nrm[c][k] = nn[c];
      }
      const bool valid = live && (loc + k) < a.count;
      if (valid) {
        acc.n_plastic += flag ? 1u : 0u;
        acc.n_fail += fail ? 1u : 0u;
        acc.max_iter = n_iter > (int)acc.max_iter ? (unsigned)n_iter : acc.max_iter;
        acc.max_resid = resid > acc.max_resid ? resid : acc.max_resid;
        if (resid != resid) acc.max_resid = resid;
        if (DIAG) {
          a.d_flag[i0 + k] = flag ? 1 : 0;
          a.d_iter[i0 + k] = n_iter;
          a.d_resid[i0 + k] = resid;
          a.d_fail[i0 + k] = fail ? 1 : 0;
        }
      }
    }

    // ---- stores: state then the 36 tangent entries (row-major j*6+i, symmetric) ---------------
    if (COMPACT && !live) continue;
#pragma unroll
    for (int c = 0; c < 6; ++c) stv<PPT>(a.sig + c * ld + i0, sig[c]);
    stv<PPT>(a.p + i0, p_new);
#pragma unroll
    for (int c = 0; c < 6; ++c) stv<PPT>(a.epsp + c * ld + i0, epsp[c]);
#pragma unroll
    for (int j = 0; j < 6; ++j) {
#pragma unroll
      for (int i = j; i < 6; ++i) {
        double v[PPT];
#pragma unroll
        for (int k = 0; k < PPT; ++k) v[k] = j2_tangent_entry(i, j, A[k], B[k], gamma[k], nrm[i][k], nrm[j][k]);
        // symmetric: each unique entry is written once, packed (sym6_packed); the boundary transposes mirror it
        stv<PPT>(a.ct + (int64_t)sym6_packed(j * 6 + i) * ld + i0, v);
      }
    }
  }
  block_reduce_stats(acc, a.stats);
}

}  // namespace dxm
