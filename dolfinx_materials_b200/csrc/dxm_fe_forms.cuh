// Fused flux / tangent -> element residual / stiffness contraction -- SURVEY.md 8(f) rank 3.
// Consumes the constitutive update's SoA outputs where they lie in HBM and produces what DOLFINx's cell kernels
// produce for  Res = dot(flux, dgrad(v)) * dx  (solvers.py:80-81; finite_strain_elastoplasticity.py:171-172,
// tests/uniaxial_tension.py:57-60)  and  Jac = qmap.derivative(Res, u, du)  (quadrature_map.py:132-158, tangent
// laid out as in quadrature_map.py:94-104), on affine simplices with a blocked Lagrange space:
//     fe[(a,r)]        = sum_q vol_q sum_j  S_rj(q) g[a,j]
//     ke[(a,r),(b,s)]  = sum_q vol_q sum_jl g[a,j] A_(rj)(sl)(q) g[b,l]        vol_q = w_q |det J|
// (kind 0: S, A from the Mandel stress / 6x6 tangent with the sqrt(2) factors undone; kind 1: PK1 and dP/dF).
// Either the element vectors / matrices are written out (MODE_ELEMENT: what MatSetValuesLocal takes), or they are
// scatter-added straight into a global vector and a CSR value array (MODE_GLOBAL): the tangent (36 | 81 doubles
// per point) then never leaves the device -- only the assembled system does.
//
// Round-2 kernel: ONE WARP PER CELL, ONE LANE PER ELEMENT-MATRIX COLUMN (b, s).
//   * the CTA (4 warps = 4 consecutive cells) stages g, vol_q g (node-major: the nqp * TDIM values of a basis function are
//     one 16-byte aligned run) and the flux / tangent of its points as plain tensors S[r][j], A[(r,j)][(s,l)] in shared
//     memory; the tangent loads of a thread are issued back to back (every SoA row is one contiguous run) and scattered
//     through a row -> tensor-position map evaluated once per CTA;
//   * a lane keeps its own g[b][:] and U_q[(r,j)] = sum_l A_q[(r,j)][(s,l)] g_q[b][l] in registers (formed once per
//     row direction), and every entry of its column is then sum_q sum_j (vol_q g_q[a][j]) U_q[(r,j)]: 12 fused
//     multiply-adds -- 15 k DFMA per P2 tetrahedron instead of the 29 k
//     DMUL + DADD and the ~130 shared-memory loads per row and point of the round-1 mapping (one thread per ROW);
//   * the row operand of an entry is carried back to the reference cell (affine simplex: one K per cell), so it is the
//     tabulated reference gradient dphi_q[a][m] -- passed to the kernel by value (constant bank), not read from shared
//     memory: ke = sum_q sum_m dphi_q[a][m] Ut_q[m], Ut_q[m] = vol_q sum_j K[m][j] U_q[j];
//   * a row of the element matrix leaves the warp as ONE predicated reduction instruction (the three columns of a node
//     are adjacent in the CSR row).  Timing-only builds (profiles/r02z_fe_forms_diag.json): the kernel is bound by the
//     issue slots / shared-memory traffic of the contraction, not by its reductions.
// Tensor cores: not used.  FP64 DMMA (mma.sync.m8n8k4.f64, the only fp64 tensor path of sm_100) would have to treat the
// gradient operator G (9 x 30) as dense although two thirds of it are structural zeros (delta_ss'): 147 kflop issued per
// cell and point set for 28 kflop of useful work, at a peak (~40 TFLOP/s) no higher than the DFMA pipe's on B200.
// Operation order == oracle/fe_forms.py (element level bit-exact; the global scatter uses fp64 atomics).
#pragma once
#include "dxm_canon.cuh"

namespace dxm {

enum FeFormMode { MODE_ELEMENT = 0, MODE_GLOBAL = 1 };

struct FeFormArgs {
  const double* coords;      // (num_nodes, 3)
  const int32_t* geom_dofs;  // (num_cells, TDIM+1)
  const int32_t* u_dofs;     // (num_cells, nd)
  const double* dphi;        // (nqp, nd, TDIM)
  const double* weights;     // (nqp)
  const double* flux;        // SoA [6|9][ld]
  const double* ct;          // SoA [21 (packed symmetric, sym6_packed) | 81][ld]
  int64_t ld, num_cells;
  int nd, nqp, kind;
  int want_vec, want_mat;
  // MODE_ELEMENT
  double* fe;  // (num_cells, ndof)
  double* ke;  // (num_cells, ndof, ndof)
  // MODE_GLOBAL
  double* b;              // (num_dofs*TDIM)
  const int64_t* rowptr;  // CSR of the blocked space, sorted columns
  const int32_t* colidx;
  double* vals;
  const uint8_t* bc;  // optional Dirichlet marker per global dof
  const double* lift;  // optional prescribed solution values on the constrained dofs: b -= A[:, bc] lift[bc]
  unsigned long long* missing;  // count of (row, col) pairs not found in the pattern
  const int32_t* off;  // optional (num_cells, nd, nd): offset of node b's column block inside node a's rows
  const int32_t* cell_list;  // optional: the cells this launch handles, num_cells = its length
};

constexpr int kFeMaxNd = 20;  // generic path: up to P3 tetrahedra
constexpr int kFeMaxQp = 8;   // Gauss points per cell the per-lane column state is sized for

DXM_HD constexpr int idx9_c(int i, int j) {
  return i == j ? i : (i == 0 && j == 1) ? 3 : (i == 1 && j == 0) ? 4 : (i == 0 && j == 2) ? 5
                  : (i == 2 && j == 0) ? 6 : (i == 1 && j == 2) ? 7 : 8;
}
DXM_HD constexpr int idx6_c(int i, int j) {
  return i == j ? i : (i + j == 1) ? 3 : (i + j == 2) ? 4 : 5;
}

// J^-1 and det J of an affine simplex, same operation order as fe_gradient_kernel / oracle.fe_forms.geometry
template <int TDIM>
DXM_HD void cell_geometry(const double* coords, const int32_t* gd, double (&K)[TDIM][TDIM], double& det) {
  double x0[TDIM], J[TDIM][TDIM];
  const double* p0 = coords + (int64_t)gd[0] * 3;
#pragma unroll
  for (int i = 0; i < TDIM; ++i) x0[i] = p0[i];
#pragma unroll
  for (int j = 0; j < TDIM; ++j) {
    const double* pj = coords + (int64_t)gd[j + 1] * 3;
#pragma unroll
    for (int i = 0; i < TDIM; ++i) J[i][j] = pj[i] - x0[i];
  }
  if (TDIM == 2) {
    det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double rdet = 1.0 / det;
    K[0][0] = J[1][1] * rdet;
    K[0][1] = -(J[0][1] * rdet);
    K[1][0] = -(J[1][0] * rdet);
    K[1][1] = J[0][0] * rdet;
  } else {
    double cf[3][3];
    cf[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    cf[0][1] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
    cf[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
    cf[1][0] = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    cf[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
    cf[1][2] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
    cf[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    cf[2][1] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
    cf[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    det = (J[0][0] * cf[0][0] + J[0][1] * cf[1][0]) + J[0][2] * cf[2][0];
    const double rdet = 1.0 / det;
#pragma unroll
    for (int i = 0; i < TDIM; ++i)
#pragma unroll
      for (int j = 0; j < TDIM; ++j) K[i][j] = cf[i][j] * rdet;
  }
}

// position of column `col` in the sorted CSR row [lo, hi); -1 if absent
__device__ __forceinline__ int64_t csr_find(const int32_t* colidx, int64_t lo, const int64_t hi, int32_t col) {
  int64_t top = hi;
  while (lo < top) {
    const int64_t mid = (lo + top) >> 1;
    if (colidx[mid] < col)
      lo = mid + 1;
    else
      top = mid;
  }
  return (lo < hi && colidx[lo] == col) ? lo : -1;
}

struct FeFormSmem {
  int cpb, np, ndof;
  size_t off_vol, off_g, off_gv, off_flux, off_ct, bytes;
};

constexpr int kFeWarps = 4;  // cells per CTA: one per warp (4 CTAs of 4 warps per SM overlap staging and contraction better than 2 of 8)

// S / A are staged as full tensors over the TDIM x TDIM gradient components: T2 = TDIM^2 entries per point for the flux,
// T2 x T2 for the tangent, entry ((r*TDIM + j), (s*TDIM + l))
inline FeFormSmem fe_form_smem(int tdim, int nd, int nqp, int kind, int mode, bool want_mat) {
  (void)kind;
  (void)mode;
  FeFormSmem s{};
  const int t2 = tdim * tdim;
  s.ndof = nd * tdim;
  s.cpb = kFeWarps;
  s.np = s.cpb * nqp;
  size_t o = 0;
  s.off_vol = o;
  o += sizeof(double) * s.np;
  s.off_g = o;
  o += sizeof(double) * (size_t)s.np * nd * tdim;
  s.off_gv = o;
  o += sizeof(double) * (size_t)s.np * nd * tdim;
  s.off_flux = o;
  o += sizeof(double) * (size_t)t2 * s.np;
  s.off_ct = o;
  o += want_mat ? sizeof(double) * (size_t)t2 * t2 * s.np : 0;
  s.bytes = o;
  return s;
}

// g[a][j] = sum_m dphi[q][a][m] K[m][j]: physical gradient of basis function a at a Gauss point (dq = dphi of that point)
template <int TDIM>
DXM_HD double fe_form_g_entry(const double* dq, const double (&K)[TDIM][TDIM], const int n, const int j) {
  double acc = dq[n * TDIM] * K[0][j];
#pragma unroll
  for (int m = 1; m < TDIM; ++m) acc = acc + dq[n * TDIM + m] * K[m][j];
  return acc;
}

// vol_q = w_q |det J| and g[a][j] of Gauss point q of `cell` (the kernel stages the same entries item by item)
template <int TDIM>
DXM_HD void fe_form_point_geometry(const FeFormArgs& a, const int64_t cell, const int q, const int nd, double& vol,
                                   double* g) {
  double K[TDIM][TDIM], det;
  cell_geometry<TDIM>(a.coords, a.geom_dofs + cell * (TDIM + 1), K, det);
  vol = a.weights[q] * fabs(det);
  const double* dq = a.dphi + (int64_t)q * nd * TDIM;
  for (int n = 0; n < nd; ++n) {
#pragma unroll
    for (int j = 0; j < TDIM; ++j) g[n * TDIM + j] = fe_form_g_entry<TDIM>(dq, K, n, j);
  }
}

// ---- staging: SoA rows of the update -> tensors over the gradient components ------------------------------------------
// tensor index pair(s) behind a Mandel index / a position of the reference's 9-vector
DXM_HD void mandel_pair(const int m, int& i, int& j) {
  i = m < 3 ? m : (m == 5 ? 1 : 0);
  j = m < 3 ? m : (m == 3 ? 1 : 2);
}
DXM_HD void vec9_pair(const int p, int& i, int& j) {
  // [11,22,33,12,21,13,31,23,32] (utils.py:173-186)
  i = p < 3 ? p : (p == 3 || p == 5 ? 0 : (p == 4 || p == 7 ? 1 : 2));
  j = p < 3 ? p : (p == 4 || p == 6 ? 0 : (p == 3 || p == 8 ? 1 : 2));
}

// Flux row `row` (Mandel 6 | reference 9-vector) of one point -> the S[r][j] entries it feeds (r, j < TDIM).
template <int TDIM>
DXM_HD void fe_stage_flux(const int kind, const int row, const double v, double* S /* [TDIM*TDIM] */) {
  constexpr double kR2 = 0.70710678118654752440;
  int i, j;
  if (kind == 1) {
    vec9_pair(row, i, j);
    if (i < TDIM && j < TDIM) S[i * TDIM + j] = v;
    return;
  }
  mandel_pair(row, i, j);
  if (i >= TDIM || j >= TDIM) return;
  if (i == j) {
    S[i * TDIM + i] = v;
  } else {
    const double w = v * kR2;
    S[i * TDIM + j] = w;
    S[j * TDIM + i] = w;
  }
}

// Tangent row `row` (packed symmetric 21 | row-major 81) of one point -> the A[(r,j)][(s,l)] entries it feeds: their
// positions in the T2 x T2 tensor (at most 8: the minor symmetries of a Mandel pair and the mirror of the symmetric 6x6
// tangent), and the factor undoing the Mandel scaling.  The kernel evaluates this map once per CTA into shared memory.
template <int TDIM>
DXM_HD void fe_tangent_map(const int kind, const int row, int16_t (&dst)[8], int& cnt, double& w) {
  constexpr double kR2 = 0.70710678118654752440;
  constexpr int T2 = TDIM * TDIM;
  int r, j, s, l;
  cnt = 0;
  w = 1.0;
  if (kind == 1) {
    vec9_pair(row / 9, r, j);
    vec9_pair(row % 9, s, l);
    if (r < TDIM && j < TDIM && s < TDIM && l < TDIM) dst[cnt++] = (int16_t)((r * TDIM + j) * T2 + s * TDIM + l);
    return;
  }
  // packed row -> Mandel pair (m1 <= m2), row-major upper triangle
  int m1 = 0, rem = row;
  while (rem >= 6 - m1) {
    rem -= 6 - m1;
    ++m1;
  }
  const int m2 = m1 + rem;
  mandel_pair(m1, r, j);
  mandel_pair(m2, s, l);
  if (r >= TDIM || j >= TDIM || s >= TDIM || l >= TDIM) return;
  const int noff = (r != j ? 1 : 0) + (s != l ? 1 : 0);
  w = noff == 0 ? 1.0 : (noff == 1 ? kR2 : 0.5);
  // every tensor entry (r,j | j,r) x (s,l | l,s) of the pair, and its mirror (the 6x6 tangent is symmetric)
  for (int t1 = 0; t1 < (r != j ? 2 : 1); ++t1)
    for (int t2 = 0; t2 < (s != l ? 2 : 1); ++t2) {
      const int rj = t1 ? j * TDIM + r : r * TDIM + j, sl = t2 ? l * TDIM + s : s * TDIM + l;
      dst[cnt++] = (int16_t)(rj * T2 + sl);
      dst[cnt++] = (int16_t)(sl * T2 + rj);
    }
}

template <int TDIM>
DXM_HD void fe_stage_tangent(const int kind, const int row, const double v, double* A /* [T2*T2] */) {
  int16_t dst[8];
  int cnt;
  double w;
  fe_tangent_map<TDIM>(kind, row, dst, cnt, w);
  const double vw = v * w;
  for (int c = 0; c < cnt; ++c) A[dst[c]] = vw;
}

// ---- one column (b, s) of one cell, from the staged arrays of the cell's nqp points -----------------------------------
// g, gv: NODE-MAJOR [nd][nqp][TDIM] (the nqp*TDIM values of one basis function are one contiguous, 16-byte aligned run:
// 128-bit shared-memory loads); S: [nqp][T2]; A: [nqp][T2*T2].  The kernel runs exactly these helpers, one lane per
// column; __host__ __device__ so that a CPU test can run them against the oracle (tests/fe_host_check.cu).

// N consecutive doubles from a 16-byte aligned address (N even on the device fast path)
template <int N>
DXM_HD void fe_load_run(const double* p, double* o) {
#ifdef __CUDA_ARCH__
  if (N % 2 == 0) {
    const double2* p2 = reinterpret_cast<const double2*>(p);
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      const double2 v = p2[i];
      o[2 * i] = v.x;
      o[2 * i + 1] = v.y;
    }
    return;
  }
#endif
#pragma unroll
  for (int i = 0; i < N; ++i) o[i] = p[i];
}

// gb[q][l] = g_q[b][l]
template <int TDIM, int NQP>
DXM_HD void fe_form_column_g(const int nqp_rt, const int nd, const int b, const double* g, double* gb /* [nqp][TDIM] */) {
  (void)nd;
  if (NQP > 0) {
    fe_load_run<(NQP > 0 ? NQP : 1) * TDIM>(g + (int64_t)b * NQP * TDIM, gb);
    return;
  }
  for (int i = 0; i < nqp_rt * TDIM; ++i) gb[i] = g[(int64_t)b * nqp_rt * TDIM + i];
}

// U[q][j] = sum_l A_q[(r,j)][(s,l)] g_q[b][l]   for the rows (., r) of the element matrix
template <int TDIM, int NQP>
DXM_HD void fe_form_column_u(const int nqp_rt, const int r, const int s, const double* gb, const double* A,
                             double* U /* [nqp][TDIM] */) {
  constexpr int T2 = TDIM * TDIM;
  const int nqp = NQP > 0 ? NQP : nqp_rt;
#pragma unroll
  for (int q = 0; q < (NQP > 0 ? NQP : kFeMaxQp); ++q) {
    if (NQP == 0 && q >= nqp) break;
    const double* Aq = A + (int64_t)q * T2 * T2;
#pragma unroll
    for (int j = 0; j < TDIM; ++j) {
      const int rj = r * TDIM + j;
      double u = Aq[rj * T2 + s * TDIM] * gb[q * TDIM];
#pragma unroll
      for (int l = 1; l < TDIM; ++l) u = fma_c(Aq[rj * T2 + s * TDIM + l], gb[q * TDIM + l], u);
      U[q * TDIM + j] = u;
    }
  }
}

// Ut[q][m] = vol_q sum_j K[m][j] U_q[j]: U carried back to the reference cell (affine simplex: g_q[a][j] = sum_m
// dphi_q[a][m] K[m][j] with ONE K per cell), so that the row operand of an entry is the tabulated reference gradient
// dphi_q[a][m] -- the same numbers for every cell of the mesh: kernel-parameter constants instead of shared-memory rows.
// Kv: K[m][j] row-major followed by |det J|; weights: the quadrature weights.
template <int TDIM, int NQP>
DXM_HD void fe_form_column_ut(const int nqp_rt, const double* Kv, const double* weights, const double* U,
                              double* Ut /* [nqp][TDIM] */) {
  const int nqp = NQP > 0 ? NQP : nqp_rt;
  const double adet = Kv[TDIM * TDIM];
#pragma unroll
  for (int q = 0; q < (NQP > 0 ? NQP : kFeMaxQp); ++q) {
    if (NQP == 0 && q >= nqp) break;
    const double vol = weights[q] * adet;
#pragma unroll
    for (int m = 0; m < TDIM; ++m) {
      double t = Kv[m * TDIM] * U[q * TDIM];
#pragma unroll
      for (int j = 1; j < TDIM; ++j) t = fma_c(Kv[m * TDIM + j], U[q * TDIM + j], t);
      Ut[q * TDIM + m] = vol * t;
    }
  }
}

// ke[(a,r),(b,s)] = sum_q sum_m dphi_q[a][m] Ut_q[m]   (q outer, m inner, one product then fused steps); dphi: the
// tabulated reference gradients [nqp][nd][TDIM]
template <int TDIM, int NQP>
DXM_HD double fe_form_entry(const int nqp_rt, const int nd, const int a, const double* dphi, const double* Ut) {
  const int nqp = NQP > 0 ? NQP : nqp_rt;
  double acc = dphi[a * TDIM] * Ut[0];
#pragma unroll
  for (int q = 0; q < (NQP > 0 ? NQP : kFeMaxQp); ++q) {
    if (NQP == 0 && q >= nqp) break;
#pragma unroll
    for (int m = 0; m < TDIM; ++m)
      if (q > 0 || m > 0) acc = fma_c(dphi[((int64_t)q * nd + a) * TDIM + m], Ut[q * TDIM + m], acc);
  }
  return acc;
}

// fe[(b,s)] = sum_q sum_j S_q[s][j] (vol_q g_q[b][j])
template <int TDIM, int NQP>
DXM_HD double fe_form_vector_entry(const int nqp_rt, const int nd, const int b, const int s, const double* gv, const double* S) {
  (void)nd;
  constexpr int T2 = TDIM * TDIM;
  const int nqp = NQP > 0 ? NQP : nqp_rt;
  double acc = 0.0;
#pragma unroll
  for (int q = 0; q < (NQP > 0 ? NQP : kFeMaxQp); ++q) {
    if (NQP == 0 && q >= nqp) break;
    const double* gbv = gv + ((int64_t)b * nqp + q) * TDIM;
#pragma unroll
    for (int j = 0; j < TDIM; ++j) {
      const double sv = S[q * T2 + s * TDIM + j];
      acc = (q == 0 && j == 0) ? sv * gbv[0] : fma_c(sv, gbv[j], acc);
    }
  }
  return acc;
}

// Resident CTAs per SM the register allocation targets: 5 (96 registers, no spills) -- 3.60 ms against 3.92 at 4 (120
// registers); 6 / 7 CTAs need the largest shared-memory carve-out, whose smaller L1 costs more than the extra warps bring
// (3.69-3.75 ms; profiles/r02w_fe_forms_ab.json).
#ifndef DXM_FE_MINB
#define DXM_FE_MINB 5
#endif
constexpr int kFeMinBlocks = DXM_FE_MINB;

// NQP > 0: Gauss points per cell at compile time (the column state stays in registers); NQP == 0: run-time count up to
// kFeMaxQp (local-memory arrays; uncommon rules)
// reference gradients and quadrature weights of the compile-time elements, passed BY VALUE: kernel parameters live in the
// constant bank, so with the row index unrolled they become immediate operands of the fused multiply-adds
template <int N, int NQ>
struct FeTab {
  double v[N > 0 ? N : 1];
  double w[NQ > 0 ? NQ : 1];
};

template <int TDIM, int ND, int NQP, int MODE>
__global__ void __launch_bounds__(32 * kFeWarps, kFeMinBlocks)
    fe_forms_kernel(const FeFormArgs a, const FeFormSmem L, const FeTab<ND * NQP * TDIM, NQP> T) {
  extern __shared__ __align__(16) double smem[];
  constexpr int T2 = TDIM * TDIM;
  constexpr int NDC = ND > 0 ? ND : kFeMaxNd;
  constexpr int NT = 32 * kFeWarps;
  double* s_g = smem + L.off_g / sizeof(double);        // [cell][nd][nqp][TDIM]
  double* s_gv = smem + L.off_gv / sizeof(double);      // [cell][nd][nqp][TDIM]   vol_q g
  double* s_flux = smem + L.off_flux / sizeof(double);  // [np][T2]
  double* s_ct = smem + L.off_ct / sizeof(double);      // [np][T2*T2]
  const int nd = ND > 0 ? ND : a.nd;
  const int ndof = nd * TDIM;
  const int nqp = NQP > 0 ? NQP : a.nqp, np = kFeWarps * nqp;
  const int nflux = a.kind == 0 ? 6 : 9, nct = a.kind == 0 ? kSym6Rows : 81;
  const int64_t c0 = (int64_t)blockIdx.x * kFeWarps;
  const int ncell = (int)min((int64_t)kFeWarps, a.num_cells - c0);
  const int npv = ncell * nqp;
  __shared__ int64_t s_cell[kFeWarps];  // the CTA's cells: consecutive, or taken from the launch's cell list
  // first CSR entry of every global row the CTA's cells touch; -1: constrained row (left alone)
  __shared__ int64_t s_rowlo[MODE == MODE_ELEMENT ? 1 : kFeWarps * kFeMaxNd * TDIM];
  __shared__ double s_K[kFeWarps][TDIM * TDIM + 1];  // J^-1 and |det J| of the CTA's cells
  // tangent row -> tensor positions (fe_tangent_map), evaluated once per CTA
  __shared__ int16_t s_dst[81][8];
  __shared__ double s_w[81];
  __shared__ int8_t s_cnt[81];
  if (threadIdx.x < kFeWarps)
    s_cell[threadIdx.x] = threadIdx.x < ncell ? (a.cell_list ? (int64_t)a.cell_list[c0 + threadIdx.x] : c0 + threadIdx.x) : 0;
  if (a.want_mat && threadIdx.x < nct) {
    int cnt;
    double w;
    int16_t dst[8];
    fe_tangent_map<TDIM>(a.kind, threadIdx.x, dst, cnt, w);
#pragma unroll
    for (int c = 0; c < 8; ++c) s_dst[threadIdx.x][c] = c < cnt ? dst[c] : (int16_t)0;
    s_w[threadIdx.x] = w;
    s_cnt[threadIdx.x] = (int8_t)cnt;
  }
  __syncthreads();

  // ---- stage: geometry -> g, vol_q g; flux / tangent rows of this CTA's points as tensors --------------------------
  // The tangent loads go first and stay in flight while the geometry is worked out: thread t owns point k = t % np and
  // the rows t / np, t / np + NT / np, ... -- every load instruction of a warp reads whole contiguous runs of SoA rows.
  const int rpp = NT / np;                                   // tangent rows per pass of the CTA
  const int tk = threadIdx.x % np, tr0 = threadIdx.x / np;   // this thread's point and first row
  const bool tlive = a.want_mat && tr0 < rpp && tk < npv;
  constexpr int kRpp = NQP > 0 ? NT / (kFeWarps * NQP) : 1;
  constexpr int kMaxPass = (81 + kRpp - 1) / kRpp;  // passes that cover the 81 rows of the finite-strain tangent
  double tv[NQP > 0 ? kMaxPass : 1];
  const double* tsrc = a.ct + s_cell[tk / nqp] * nqp + (tk % nqp);
  if (NQP > 0) {
#pragma unroll
    for (int m = 0; m < kMaxPass; ++m) {
      const int row = tr0 + m * rpp;
      tv[m] = (tlive && row < nct) ? __ldcs(tsrc + (int64_t)row * a.ld) : 0.0;
    }
  }
  if (threadIdx.x < ncell) {
    double K[TDIM][TDIM], det;
    cell_geometry<TDIM>(a.coords, a.geom_dofs + s_cell[threadIdx.x] * (TDIM + 1), K, det);
#pragma unroll
    for (int i = 0; i < TDIM; ++i)
#pragma unroll
      for (int j = 0; j < TDIM; ++j) s_K[threadIdx.x][i * TDIM + j] = K[i][j];
    s_K[threadIdx.x][TDIM * TDIM] = fabs(det);
  }
  for (int i = threadIdx.x; i < nflux * np; i += NT) {
    const int row = i / np, k = i - row * np;
    if (k < npv) {
      const int lc = k / nqp;
      fe_stage_flux<TDIM>(a.kind, row, __ldcs(a.flux + (int64_t)row * a.ld + s_cell[lc] * nqp + (k - lc * nqp)),
                          s_flux + (int64_t)k * T2);
    }
  }
  if (a.want_mat && MODE != MODE_ELEMENT) {
    for (int i = threadIdx.x; i < ncell * ndof; i += NT) {
      const int lc = i / ndof, row = i - lc * ndof;
      const int64_t grow = (int64_t)a.u_dofs[s_cell[lc] * nd + row / TDIM] * TDIM + row % TDIM;
      s_rowlo[lc * kFeMaxNd * TDIM + row] = (a.bc && a.bc[grow]) ? -1 : a.rowptr[grow];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < npv * nd; i += NT) {  // one (point, basis function) per item
    const int pt = i / nd, n = i - pt * nd;
    const int lc = pt / nqp, q = pt - lc * nqp;
    double K[TDIM][TDIM];
#pragma unroll
    for (int ii = 0; ii < TDIM; ++ii)
#pragma unroll
      for (int jj = 0; jj < TDIM; ++jj) K[ii][jj] = s_K[lc][ii * TDIM + jj];
    const double vol = a.weights[q] * s_K[lc][TDIM * TDIM];
    const double* dq = a.dphi + (int64_t)q * nd * TDIM;
    const int64_t o = (((int64_t)lc * nd + n) * nqp + q) * TDIM;
#pragma unroll
    for (int j = 0; j < TDIM; ++j) {
      const double gk = fe_form_g_entry<TDIM>(dq, K, n, j);
      s_g[o + j] = gk;
      s_gv[o + j] = vol * gk;
    }
  }
  if (a.want_mat) {
    if (NQP > 0) {
      double* dstA = s_ct + (int64_t)tk * T2 * T2;
#pragma unroll
      for (int m = 0; m < kMaxPass; ++m) {
        const int row = tr0 + m * rpp;
        if (tlive && row < nct) {
          const double vw = tv[m] * s_w[row];
          const int cnt = s_cnt[row];
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (c < cnt) dstA[s_dst[row][c]] = vw;
        }
      }
    } else {
      for (int i = threadIdx.x; i < nct * np; i += NT) {
        const int row = i / np, k = i - row * np;
        if (k < npv) {
          const int lc = k / nqp;
          fe_stage_tangent<TDIM>(a.kind, row, __ldcs(a.ct + (int64_t)row * a.ld + s_cell[lc] * nqp + (k - lc * nqp)),
                                 s_ct + (int64_t)k * T2 * T2);
        }
      }
    }
  }
  __syncthreads();

  const int lc = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lc >= ncell) return;
  const int64_t cell = s_cell[lc];
  const double* g = s_g + (int64_t)lc * nqp * nd * TDIM;
  const double* gv = s_gv + (int64_t)lc * nqp * nd * TDIM;
  const double* S = s_flux + (int64_t)lc * nqp * T2;
  const double* A = s_ct + (int64_t)lc * nqp * T2 * T2;
  const int32_t* ud = a.u_dofs + cell * nd;
  const int64_t* rowlo = s_rowlo + (MODE == MODE_ELEMENT ? 0 : lc * kFeMaxNd * TDIM);
  const double* Kv = s_K[lc];                                   // K[m][j] row-major, |det J|
  const double* tab = ND > 0 ? T.v : a.dphi;                    // reference gradients: constants | global memory
  const double* wts = (ND > 0 && NQP > 0) ? T.w : a.weights;

  for (int cb = 0; cb < ndof; cb += 32) {  // columns in chunks of a warp (one chunk up to 10 nodes in 3-D)
    const int col = cb + lane;
    const bool live = col < ndof;
    const int b = live ? col / TDIM : 0, s = live ? col - b * TDIM : 0;
    double gb[kFeMaxQp * TDIM], U[kFeMaxQp * TDIM], Ut[kFeMaxQp * TDIM];
    fe_form_column_g<TDIM, NQP>(nqp, nd, b, g, gb);

    if (MODE == MODE_ELEMENT) {
      if (live && a.want_vec) a.fe[cell * ndof + col] = fe_form_vector_entry<TDIM, NQP>(nqp, nd, b, s, gv, S);
      if (!a.want_mat) continue;
#pragma unroll 1
      for (int r = 0; r < TDIM; ++r) {
        fe_form_column_u<TDIM, NQP>(nqp, r, s, gb, A, U);
        fe_form_column_ut<TDIM, NQP>(nqp, Kv, wts, U, Ut);
#pragma unroll
        for (int an = 0; an < NDC; ++an) {
          if (ND == 0 && an >= nd) break;
          const double v = fe_form_entry<TDIM, NQP>(nqp, nd, an, tab, Ut);
          if (live) __stcs(a.ke + (cell * ndof + an * TDIM + r) * ndof + col, v);  // a row is one contiguous line
        }
      }
      continue;
    }

    const int64_t gcol = live ? (int64_t)ud[b] * TDIM + s : 0;
    const bool col_bc = live && a.bc && a.bc[gcol];
    if (live && a.want_vec && !col_bc) atomicAdd(a.b + gcol, fe_form_vector_entry<TDIM, NQP>(nqp, nd, b, s, gv, S));
    if (!a.want_mat) continue;
    const double lift = (col_bc && a.lift) ? a.lift[gcol] : 0.0;
    const bool any_lift = a.lift && __any_sync(0xffffffffu, col_bc);
    const bool writer = live && !col_bc;
    unsigned miss = 0;
    if (a.off && !any_lift) {
      // node-blocked pattern, no constrained column in this cell (all but the cells on a Dirichlet boundary): the offset
      // of (node a, node b) inside the rows of node a was found once (fe_offsets_kernel) and the TDIM columns of node b
      // are consecutive there -- a row of the element matrix leaves the warp as one predicated reduction instruction
      int32_t offs[NDC];
#pragma unroll
      for (int an = 0; an < NDC; ++an) {
        if (ND == 0 && an >= nd) break;
        offs[an] = writer ? a.off[(cell * nd + an) * nd + b] : -1;
      }
      // (r rolled on purpose: unrolled, the compiler keeps the r-invariant vol g rows of all nodes live across the three
      // passes and spills them to local memory instead of reading shared memory again)
#pragma unroll 1
      for (int r = 0; r < TDIM; ++r) {
        fe_form_column_u<TDIM, NQP>(nqp, r, s, gb, A, U);
        fe_form_column_ut<TDIM, NQP>(nqp, Kv, wts, U, Ut);
#pragma unroll
        for (int an = 0; an < NDC; ++an) {
          if (ND == 0 && an >= nd) break;
          const int64_t lo = rowlo[an * TDIM + r];  // < 0: constrained row, untouched (unit diagonal set by the host API)
          const double v = fe_form_entry<TDIM, NQP>(nqp, nd, an, tab, Ut);
          const bool row_ok = writer && lo >= 0;
          if (row_ok && offs[an] >= 0) atomicAdd(a.vals + (lo + offs[an] + s), v);  // result unused: a RED
          miss += (row_ok && offs[an] < 0) ? 1u : 0u;
        }
      }
    } else {
      // cells with a constrained column (lifting) and patterns without an offset table (binary search of the column in
      // the sorted CSR row, per entry): rolled loops
#pragma unroll 1
      for (int r = 0; r < TDIM; ++r) {
        fe_form_column_u<TDIM, NQP>(nqp, r, s, gb, A, U);
        fe_form_column_ut<TDIM, NQP>(nqp, Kv, wts, U, Ut);
#pragma unroll 1
        for (int an = 0; an < nd; ++an) {
          const int64_t lo = rowlo[an * TDIM + r];
          if (lo < 0) continue;
          const double v = fe_form_entry<TDIM, NQP>(nqp, nd, an, a.dphi, Ut);  // rolled: the table from global memory
          if (any_lift) {
            // constrained columns move to the right-hand side (apply_lifting): b[row] -= sum_bc K[row][col] lift[col]
            double lf = col_bc ? v * lift : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) lf += __shfl_xor_sync(0xffffffffu, lf, o);
            if (lane == 0 && lf != 0.0) atomicAdd(a.b + (int64_t)ud[an] * TDIM + r, -lf);
          }
          if (writer) {
            int64_t p;
            if (a.off) {
              const int32_t o = a.off[(cell * nd + an) * nd + b];
              p = o >= 0 ? lo + o + s : -1;
            } else {
              const int64_t grow = (int64_t)ud[an] * TDIM + r;
              p = csr_find(a.colidx, lo, a.rowptr[grow + 1], (int32_t)gcol);
            }
            if (p >= 0)
              atomicAdd(a.vals + p, v);
            else
              ++miss;
          }
        }
      }
    }
    if (miss) atomicAdd(a.missing, (unsigned long long)miss);
  }
}

// ---- residual only: one thread per (cell, basis function) -----------------------------------------------------------
// fe[(b, s)] = sum_q sum_j S_q[s][j] (vol_q g_q[b][j]) needs no staging: 12 fused multiply-adds per component on values
// the threads of a cell share through L1.  The warp-per-cell kernel above spends 0.86 ms on 663 552 P2 tetrahedra for this
// (three block barriers and three dependent load phases per 4 cells); this one is bound by its 191 MB of flux reads.
// Same operations in the same order as fe_form_vector_entry (bit-identical; tests/fe_host_check.cu compares the two).
template <int TDIM, int ND, int NQP>
DXM_HD void fe_vector_node(const FeFormArgs& a, const int64_t cell, const int n, double (&acc)[TDIM]) {
  constexpr int T2 = TDIM * TDIM;
  const int nflux = a.kind == 0 ? 6 : 9;
  double K[TDIM][TDIM], det;
  cell_geometry<TDIM>(a.coords, a.geom_dofs + cell * (TDIM + 1), K, det);
  const double adet = fabs(det);
#pragma unroll
  for (int q = 0; q < NQP; ++q) {
    const double vol = a.weights[q] * adet;
    const double* dq = a.dphi + (int64_t)q * ND * TDIM;
    double gv[TDIM], S[T2];
#pragma unroll
    for (int j = 0; j < TDIM; ++j) gv[j] = vol * fe_form_g_entry<TDIM>(dq, K, n, j);
#pragma unroll
    for (int i = 0; i < T2; ++i) S[i] = 0.0;
#pragma unroll
    for (int row = 0; row < 9; ++row)
      if (row < nflux) fe_stage_flux<TDIM>(a.kind, row, a.flux[(int64_t)row * a.ld + cell * NQP + q], S);
#pragma unroll
    for (int s = 0; s < TDIM; ++s)
#pragma unroll
      for (int j = 0; j < TDIM; ++j) {
        const double sv = S[s * TDIM + j];
        acc[s] = (q == 0 && j == 0) ? sv * gv[0] : fma_c(sv, gv[j], acc[s]);
      }
  }
}

constexpr int kFeVecBlock = 256;
template <int TDIM, int ND, int NQP, int MODE>
__global__ void __launch_bounds__(kFeVecBlock) fe_vector_kernel(const FeFormArgs a) {
  const int64_t t = (int64_t)blockIdx.x * kFeVecBlock + threadIdx.x;
  const int64_t ci = t / ND;
  const int n = (int)(t - ci * ND);
  if (ci >= a.num_cells) return;
  const int64_t cell = a.cell_list ? (int64_t)a.cell_list[ci] : ci;
  double acc[TDIM];
  fe_vector_node<TDIM, ND, NQP>(a, cell, n, acc);
  if (MODE == MODE_ELEMENT) {
#pragma unroll
    for (int s = 0; s < TDIM; ++s) a.fe[(cell * ND + n) * TDIM + s] = acc[s];
    return;
  }
  const int64_t g0 = (int64_t)a.u_dofs[cell * ND + n] * TDIM;
#pragma unroll
  for (int s = 0; s < TDIM; ++s)
    if (!(a.bc && a.bc[g0 + s])) atomicAdd(a.b + g0 + s, acc[s]);
}

// ---- atomic-free assembly: element matrices -> CSR rows, node by node ---------------------------------------------------
// The L2 retires ~90 G scalar fp64 reductions per second however they are grouped, which bounds the atomic scatter of the
// 900 entries of a P2 tetrahedron at 6.5 ms for 663 k cells (profiles/r02g_fe_forms_assembly_experiments.json).  Instead:
// fe_forms_kernel<MODE_ELEMENT> writes every element matrix as full contiguous lines, and ONE WARP PER MESH NODE then sums
// the rows (a, r) of the cells around its node into a shared-memory image of the node's TDIM CSR rows -- cell after
// cell in a fixed order, each lane owning distinct columns, so no atomics and bit-reproducible sums -- applies the
// Dirichlet rows / columns (apply_lifting) and writes the rows out as whole lines.  Everything streams: element matrices
// written once and read once, CSR values written once (no memset).
struct FeGatherArgs {
  const double* ke;        // (num_cells, ndof, ndof)
  const double* fe;        // (num_cells, ndof)
  const int64_t* nc_ptr;   // node -> [cells around it], CSR over the nodes
  const int32_t* nc_cell;
  const uint8_t* nc_loc;   // local index of the node in that cell
  const int32_t* off;      // (num_cells, nd, nd): offset of node b's column block inside node a's rows
  const int64_t* rowptr;
  const int32_t* colidx;
  double* vals;
  double* rhs;
  const uint8_t* bc;
  const double* lift;
  int64_t num_nodes;
  int nd, want_vec, maxlen;
};

constexpr int kGatherWarps = 8;
constexpr int kGatherAhead = 4;  // cells whose element-matrix rows are requested before the first is accumulated

template <int TDIM>
__global__ void __launch_bounds__(32 * kGatherWarps) fe_gather_kernel(const FeGatherArgs a) {
  extern __shared__ double smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * kGatherWarps + w;
  if (n >= a.num_nodes) return;
  double* buf = smem + (size_t)w * TDIM * a.maxlen;  // [TDIM][maxlen]: the node's rows
  const int nd = a.nd, ndof = nd * TDIM;
  const int64_t lo0 = a.rowptr[n * TDIM];
  const int len = (int)(a.rowptr[n * TDIM + 1] - lo0);  // node-blocked pattern: the TDIM rows of a node have one structure
#pragma unroll
  for (int r = 0; r < TDIM; ++r)
    for (int k = lane; k < len; k += 32) buf[r * a.maxlen + k] = 0.0;
  __syncwarp();
  double facc = 0.0;  // lane r < TDIM: vector entry of row r
  const int64_t e0 = a.nc_ptr[n], e1 = a.nc_ptr[n + 1];
  if (ndof <= 32) {
    // the usual elements: a row of the element matrix is one warp access; the loads of kGatherAhead cells are in flight
    // before the first of them is accumulated (cell after cell: the summation order stays fixed)
    const int col = lane;
    const bool live = col < ndof;
    const int b = live ? col / TDIM : 0, s = live ? col - b * TDIM : 0;
    for (int64_t e = e0; e < e1; e += kGatherAhead) {
      double v[kGatherAhead][TDIM], f[kGatherAhead];
      int o[kGatherAhead];
#pragma unroll
      for (int u = 0; u < kGatherAhead; ++u) {
        const bool on = e + u < e1;
        const int64_t c = on ? a.nc_cell[e + u] : 0;
        const int la = on ? a.nc_loc[e + u] : 0;
        o[u] = (on && live) ? a.off[(c * nd + la) * nd + b] + s : -1;
        const double* rows = a.ke + (c * ndof + (int64_t)la * TDIM) * ndof;
#pragma unroll
        for (int r = 0; r < TDIM; ++r) v[u][r] = o[u] >= 0 ? __ldcs(rows + (int64_t)r * ndof + col) : 0.0;
        f[u] = (on && a.want_vec && lane < TDIM) ? __ldcs(a.fe + c * ndof + la * TDIM + lane) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < kGatherAhead; ++u) {
        if (o[u] >= 0) {
#pragma unroll
          for (int r = 0; r < TDIM; ++r) buf[r * a.maxlen + o[u]] += v[u][r];
        }
        facc += f[u];
        __syncwarp();  // the next cell's lanes may own the same columns
      }
    }
  } else {
    for (int64_t e = e0; e < e1; ++e) {
      const int64_t c = a.nc_cell[e];
      const int la = a.nc_loc[e];
      const double* rows = a.ke + (c * ndof + (int64_t)la * TDIM) * ndof;
      for (int cb = 0; cb < ndof; cb += 32) {
        const int col = cb + lane;
        if (col < ndof) {
          const int b = col / TDIM, s = col - b * TDIM;
          const int o = a.off[(c * nd + la) * nd + b] + s;
#pragma unroll
          for (int r = 0; r < TDIM; ++r) buf[r * a.maxlen + o] += __ldcs(rows + (int64_t)r * ndof + col);
        }
      }
      if (a.want_vec && lane < TDIM) facc += __ldcs(a.fe + c * ndof + la * TDIM + lane);
      __syncwarp();
    }
  }
#pragma unroll
  for (int r = 0; r < TDIM; ++r) {
    const int64_t i = n * TDIM + r, lo = a.rowptr[i];
    const bool row_bc = a.bc && a.bc[i];
    double lifted = 0.0;
    for (int k = lane; k < len; k += 32) {
      double v = buf[r * a.maxlen + k];
      if (a.bc) {
        const int32_t j = a.colidx[lo + k];
        if (row_bc) {
          v = 0.0;  // constrained row: empty here, unit diagonal set by fe_bc_diag_kernel
        } else if (a.bc[j]) {
          if (a.lift) lifted += v * a.lift[j];  // constrained column: moved to the right-hand side (apply_lifting)
          v = 0.0;
        }
      }
      __stcs(a.vals + lo + k, v);
    }
    if (a.want_vec) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lifted += __shfl_xor_sync(0xffffffffu, lifted, o);
      const double f = __shfl_sync(0xffffffffu, facc, r);
      if (lane == 0) a.rhs[i] = row_bc ? 0.0 : f - lifted;
    }
  }
}

// One-off per (mesh, pattern): offset of node b's column block inside the rows of node a, for every cell.  Valid
// only for node-blocked patterns (the TDIM rows of a node share one column structure made of whole TDIM-blocks,
// which is what DOLFINx builds for a blocked space); `bad` counts violations -> the caller keeps the search path.
__global__ void fe_offsets_kernel(const int32_t* u_dofs, int64_t num_cells, int nd, int tdim, const int64_t* rowptr,
                                  const int32_t* colidx, int32_t* off, unsigned long long* bad) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= num_cells * nd * nd) return;
  const int64_t c = i / (nd * nd);
  const int an = (int)((i / nd) % nd), b = (int)(i % nd);
  const int32_t* ud = u_dofs + c * nd;
  const int64_t row0 = (int64_t)ud[an] * tdim;
  const int32_t col0 = ud[b] * tdim;
  const int64_t lo = rowptr[row0], hi = rowptr[row0 + 1];
  const int64_t pos = csr_find(colidx, lo, hi, col0);
  bool ok = pos >= 0 && pos + tdim <= hi;
  if (ok)
    for (int r = 0; r < tdim && ok; ++r) {
      const int64_t lr = rowptr[row0 + r];
      ok = rowptr[row0 + r + 1] - lr == hi - lo;
      for (int s = 0; s < tdim && ok; ++s) ok = colidx[lr + (pos - lo) + s] == col0 + s;
    }
  off[i] = ok ? (int32_t)(pos - lo) : -1;
  if (!ok) atomicAdd(bad, 1ull);
}

// unit diagonal on constrained rows (assemble_matrix(..., bcs) convention); rhs = prescribed value (set_bc)
__global__ void fe_bc_diag_kernel(const uint8_t* bc, const int64_t* rowptr, const int32_t* colidx, double* vals,
                                  int64_t nrows, const double* lift, double* rhs) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nrows || !bc[i]) return;
  if (rhs) rhs[i] = lift ? lift[i] : 0.0;
  if (!vals) return;
  const int64_t lo = rowptr[i], hi = rowptr[i + 1];
  if (lo >= hi) return;
  const int64_t pos = csr_find(colidx, lo, hi, (int32_t)i);
  if (pos >= 0) vals[pos] = 1.0;
}

}  // namespace dxm
