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
// Operation order == oracle/fe_forms.py (element level bit-exact; the global scatter uses fp64 atomics).
//
// Mapping: one thread per element-matrix row (cell, a, r); CPB = 256 / (ND*TDIM) cells per CTA.  The CTA stages
// the flux and tangent rows of its CPB*nqp consecutive Gauss points in shared memory with coalesced loads (each
// SoA row contributes one contiguous run), computes g once per cell, and in MODE_ELEMENT transposes the rows through
// shared memory so that the (num_cells, ndof, ndof) output is written in full contiguous lines.
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
};

constexpr int kFeMaxNd = 20;  // generic path: up to P3 tetrahedra

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
  size_t off_vol, off_g, off_flux, off_ct, off_out, bytes;
};

inline FeFormSmem fe_form_smem(int tdim, int nd, int nqp, int kind, int mode, bool want_mat) {
  FeFormSmem s{};
  s.ndof = nd * tdim;
  s.cpb = 256 / s.ndof;
  if (s.cpb < 1) s.cpb = 1;
  s.np = s.cpb * nqp;
  const int nflux = kind == 0 ? 6 : 9, nct = kind == 0 ? kSym6Rows : 81;
  size_t o = 0;
  s.off_vol = o;
  o += sizeof(double) * s.np;
  s.off_g = o;
  o += sizeof(double) * (size_t)s.np * nd * tdim;
  s.off_flux = o;
  o += sizeof(double) * (size_t)nflux * s.np;
  s.off_ct = o;
  o += want_mat ? sizeof(double) * (size_t)nct * s.np : 0;
  s.off_out = o;
  if (mode == MODE_ELEMENT && want_mat) o += sizeof(double) * (size_t)s.cpb * s.ndof * (s.ndof + 1);
  s.bytes = o;
  return s;
}

// vol_q = w_q |det J| and g[a][j] = sum_m dphi[q][a][m] K[m][j] of Gauss point q of `cell` (the kernel's staging step)
template <int TDIM>
DXM_HD void fe_form_point_geometry(const FeFormArgs& a, const int64_t cell, const int q, const int nd, double& vol,
                                   double* g) {
  double K[TDIM][TDIM], det;
  cell_geometry<TDIM>(a.coords, a.geom_dofs + cell * (TDIM + 1), K, det);
  vol = a.weights[q] * fabs(det);
  const double* dq = a.dphi + (int64_t)q * nd * TDIM;
  for (int n = 0; n < nd; ++n) {
#pragma unroll
    for (int j = 0; j < TDIM; ++j) {
      double acc = dq[n * TDIM] * K[0][j];
#pragma unroll
      for (int m = 1; m < TDIM; ++m) acc = acc + dq[n * TDIM + m] * K[m][j];
      g[n * TDIM + j] = acc;
    }
  }
}

// Row (a, r) of the element vector / matrix of local cell `lc` from the staged arrays: vol [np], g [np][nd][TDIM],
// flux [nflux][np], ct [nct][np] (shared memory in the kernel).  __host__ __device__ like the point routines of the
// constitutive kernels, so that a CPU test can run it against the oracle (tests/fe_host_check.cu).
template <int TDIM, int ND>
DXM_HD void fe_form_row(const int kind, const bool want_mat, const int nqp, const int nd, const int np, const int lc,
                        const int an, const int r, const double* s_vol, const double* s_g, const double* s_flux,
                        const double* s_ct, double& fe, double* acc) {
  constexpr double kR2 = 0.70710678118654752440;
  constexpr int NDC = ND > 0 ? ND : kFeMaxNd;
  for (int q = 0; q < nqp; ++q) {
    const int pt = lc * nqp + q;
    const double vol = s_vol[pt];
    const double* g = s_g + (int64_t)pt * nd * TDIM;
    double ga[TDIM];
#pragma unroll
    for (int j = 0; j < TDIM; ++j) ga[j] = g[an * TDIM + j];
    // residual row
    {
      double t = 0.0;
#pragma unroll
      for (int j = 0; j < TDIM; ++j) {
        double S;
        if (kind == 1) {
          S = s_flux[idx9_c(r, j) * np + pt];
        } else {
          S = s_flux[idx6_c(r, j) * np + pt];
          if (r != j) S = S * kR2;
        }
        t = j == 0 ? S * ga[0] : t + S * ga[j];
      }
      fe = q == 0 ? vol * t : fe + vol * t;
    }
    if (!want_mat) continue;
    // W[s][l] = sum_j g[a][j] A(rj, sl)
    double W[TDIM][TDIM];
#pragma unroll
    for (int s = 0; s < TDIM; ++s)
#pragma unroll
      for (int l = 0; l < TDIM; ++l) {
        double w = 0.0;
#pragma unroll
        for (int j = 0; j < TDIM; ++j) {
          double A;
          if (kind == 1) {
            A = s_ct[(idx9_c(r, j) * 9 + idx9_c(s, l)) * np + pt];
          } else {
            A = s_ct[sym6_packed(idx6_c(r, j) * 6 + idx6_c(s, l)) * np + pt];
            const int noff = (r != j ? 1 : 0) + (s != l ? 1 : 0);
            if (noff == 1) A = A * kR2;
            if (noff == 2) A = A * 0.5;
          }
          w = j == 0 ? ga[0] * A : w + ga[j] * A;
        }
        W[s][l] = w;
      }
#pragma unroll
    for (int b = 0; b < NDC; ++b) {
      if (ND == 0 && b >= nd) break;
#pragma unroll
      for (int s = 0; s < TDIM; ++s) {
        double t2 = W[s][0] * g[b * TDIM];
#pragma unroll
        for (int l = 1; l < TDIM; ++l) t2 = t2 + W[s][l] * g[b * TDIM + l];
        acc[b * TDIM + s] = q == 0 ? vol * t2 : acc[b * TDIM + s] + vol * t2;
      }
    }
  }
}

template <int TDIM, int ND, int MODE>
__global__ void __launch_bounds__(256) fe_forms_kernel(const FeFormArgs a, const FeFormSmem L) {
  extern __shared__ double smem[];
  double* s_vol = smem + L.off_vol / sizeof(double);    // [np]
  double* s_g = smem + L.off_g / sizeof(double);        // [np][nd][TDIM]
  double* s_flux = smem + L.off_flux / sizeof(double);  // [nflux][np]
  double* s_ct = smem + L.off_ct / sizeof(double);      // [nct][np]
  double* s_out = smem + L.off_out / sizeof(double);    // [cpb*ndof][ndof+1]
  const int nd = ND > 0 ? ND : a.nd;
  const int ndof = nd * TDIM;
  const int cpb = L.cpb, nqp = a.nqp, np = L.np;
  const int nflux = a.kind == 0 ? 6 : 9, nct = a.kind == 0 ? kSym6Rows : 81;
  const int64_t c0 = (int64_t)blockIdx.x * cpb;
  const int ncell = (int)min((int64_t)cpb, a.num_cells - c0);
  const int npv = ncell * nqp;
  const int64_t p0 = c0 * nqp;

  // ---- stage: geometry -> vol_q, g[q][a][j]; flux / tangent rows of this CTA's points -------------------
  for (int i = threadIdx.x; i < npv; i += blockDim.x) {
    const int lc = i / nqp, q = i - lc * nqp;
    fe_form_point_geometry<TDIM>(a, c0 + lc, q, nd, s_vol[i], s_g + (int64_t)i * nd * TDIM);
  }
  for (int i = threadIdx.x; i < nflux * np; i += blockDim.x) {
    const int row = i / np, k = i - row * np;
    if (k < npv) s_flux[i] = __ldcs(a.flux + (int64_t)row * a.ld + p0 + k);
  }
  if (a.want_mat) {
    for (int i = threadIdx.x; i < nct * np; i += blockDim.x) {
      const int row = i / np, k = i - row * np;
      if (k < npv) s_ct[i] = __ldcs(a.ct + (int64_t)row * a.ld + p0 + k);
    }
  }
  __syncthreads();

  const int lc = threadIdx.x / ndof;
  const int row = threadIdx.x - lc * ndof;
  const int an = row / TDIM, r = row - an * TDIM;
  const bool live = lc < ncell;
  constexpr int NDC = ND > 0 ? ND : kFeMaxNd;
  double fe = 0.0;
  double acc[NDC * TDIM];
  if (live) fe_form_row<TDIM, ND>(a.kind, a.want_mat != 0, nqp, nd, np, lc, an, r, s_vol, s_g, s_flux, s_ct, fe, acc);

  if (MODE == MODE_ELEMENT) {
    if (live && a.want_vec) a.fe[(c0 + lc) * ndof + row] = fe;
    if (a.want_mat) {
      // transpose through shared memory: thread rows -> contiguous (cell, row, col) lines
      if (live) {
        double* o = s_out + (int64_t)threadIdx.x * (ndof + 1);
#pragma unroll
        for (int k = 0; k < NDC * TDIM; ++k) {
          if (ND == 0 && k >= ndof) break;
          o[k] = acc[k];
        }
      }
      __syncthreads();
      const int total = ncell * ndof * ndof;
      double* dst = a.ke + c0 * ndof * ndof;
      for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int rw = i / ndof, cl = i - rw * ndof;
        __stcs(dst + i, s_out[(int64_t)rw * (ndof + 1) + cl]);
      }
    }
  } else {
    if (!live) return;
    const int32_t* ud = a.u_dofs + (c0 + lc) * nd;
    const int64_t grow = (int64_t)ud[an] * TDIM + r;
    if (a.bc && a.bc[grow]) return;  // constrained row: untouched (unit diagonal set by the host API)
    if (a.want_vec) atomicAdd(a.b + grow, fe);
    if (!a.want_mat) return;
    const int64_t lo = a.rowptr[grow], hi = a.rowptr[grow + 1];
    unsigned miss = 0;
    double lifted = 0.0;
#pragma unroll
    for (int b = 0; b < NDC; ++b) {
      if (ND == 0 && b >= nd) break;
      const int32_t cbase = ud[b] * TDIM;
      // node-blocked patterns: the offset of (node a, node b) inside the row was found once (fe_offsets_kernel)
      int64_t pos = a.off ? (a.off[((c0 + lc) * nd + an) * nd + b] >= 0 ? lo + a.off[((c0 + lc) * nd + an) * nd + b] : -1)
                          : csr_find(a.colidx, lo, hi, cbase);
#pragma unroll
      for (int s = 0; s < TDIM; ++s) {
        const int32_t gcol = cbase + s;
        if (s > 0) {
          if (a.off)
            pos = pos >= 0 ? pos + 1 : -1;
          // blocked pattern: the columns of one node are consecutive; fall back to a search otherwise
          else if (pos >= 0 && pos + 1 < hi && a.colidx[pos + 1] == gcol)
            pos = pos + 1;
          else
            pos = csr_find(a.colidx, lo, hi, gcol);
        }
        if (a.bc && a.bc[gcol]) {
          // constrained column: moved to the right-hand side (apply_lifting)
          if (a.lift) lifted += acc[b * TDIM + s] * a.lift[gcol];
          continue;
        }
        if (pos < 0) {
          ++miss;
          continue;
        }
        atomicAdd(a.vals + pos, acc[b * TDIM + s]);
      }
    }
    if (a.lift && lifted != 0.0) atomicAdd(a.b + grow, -lifted);
    if (miss) atomicAdd(a.missing, (unsigned long long)miss);
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
