// GPU evaluation of the displacement-gradient expressions that feed Material.integrate -- SURVEY.md 8(f)
// rank 2.  Replaces, for affine simplex meshes with a blocked Lagrange displacement space,
// QuadratureExpression.eval (dolfinx_materials/quadrature_function.py:45-51: fem.Expression.eval over all
// cells + scatter) and the gather of quadrature_map.py:251-253, 305-313: only the displacement vector
// crosses PCIe, the (n, 6|9) gradient array never exists on the host.  Results go straight into the
// material's SoA gradient buffer (point = num_qp*cell + q, quadrature_map.py:255-260).
//   kind 0: Mandel vector of sym(grad u)            (utils.py:146-165; 2-D pads with zeros)
//   kind 1: [11,22,33,12,21,13,31,23,32] of I+grad u (utils.py:168-190)
// Operation order == oracle/fe_gradient.py.  One thread per cell: J^-1 once, nodal displacements held in
// registers (ND compile-time for P1/P2 simplices), all quadrature points of the cell written by the thread.
#pragma once
#include "dxm_canon.cuh"

namespace dxm {

struct FeGradArgs {
  const double* coords;      // (num_nodes, 3)
  const int32_t* geom_dofs;  // (num_cells, TDIM+1)
  const int32_t* u_dofs;     // (num_cells, nd)
  const double* u;           // (num_dofs * TDIM), blocked
  const double* dphi;        // (nqp, nd, TDIM)
  double* out;               // SoA gradient buffer [ncomp][ld]
  int64_t ld;
  int64_t num_cells;
  int nd, nqp, kind;
};

constexpr int kFeMaxTab = 4 * 10 * 3 * 4;  // doubles of tabulated gradients staged in shared memory

// One cell: J^-1, nodal displacements, every quadrature point of the cell.  __host__ __device__ so that a CPU test can
// run the very code the kernel runs per cell against the oracle (tests/fe_host_check.cu) -- the product only ever calls
// it from the kernel below.  `dphi`: the tabulated gradients (shared memory in the kernel).
template <int TDIM, int ND>
DXM_HD void fe_gradient_cell(const FeGradArgs& a, const double* dphi, const int nd, const int64_t c) {
  constexpr double kR2 = 0.70710678118654752440;

  // affine geometry: J[i][j] = x_{j+1}[i] - x_0[i]
  double x0[TDIM], J[TDIM][TDIM], K[TDIM][TDIM];
  const int32_t* gd = a.geom_dofs + c * (TDIM + 1);
  {
    const double* p0 = a.coords + (int64_t)gd[0] * 3;
#pragma unroll
    for (int i = 0; i < TDIM; ++i) x0[i] = p0[i];
#pragma unroll
    for (int j = 0; j < TDIM; ++j) {
      const double* pj = a.coords + (int64_t)gd[j + 1] * 3;
#pragma unroll
      for (int i = 0; i < TDIM; ++i) J[i][j] = pj[i] - x0[i];
    }
  }
  if (TDIM == 2) {
    const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
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
    const double det = (J[0][0] * cf[0][0] + J[0][1] * cf[1][0]) + J[0][2] * cf[2][0];
    const double rdet = 1.0 / det;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) K[i][j] = cf[i][j] * rdet;
  }

  const int32_t* ud = a.u_dofs + c * nd;
  constexpr int NDR = ND > 0 ? ND : 1;
  double ua[NDR][TDIM];
  if (ND > 0) {
#pragma unroll
    for (int n = 0; n < ND; ++n) {
      const double* pu = a.u + (int64_t)ud[n] * TDIM;
#pragma unroll
      for (int r = 0; r < TDIM; ++r) ua[n][r] = pu[r];
    }
  }

  for (int q = 0; q < a.nqp; ++q) {
    double H[TDIM][TDIM];
    const double* dq = dphi + (int64_t)q * nd * TDIM;
    if (ND > 0) {
#pragma unroll
      for (int r = 0; r < TDIM; ++r)
#pragma unroll
        for (int j = 0; j < TDIM; ++j) H[r][j] = ua[0][r] * dq[j];
#pragma unroll
      for (int n = 1; n < ND; ++n)
#pragma unroll
        for (int r = 0; r < TDIM; ++r)
#pragma unroll
          for (int j = 0; j < TDIM; ++j) H[r][j] = H[r][j] + ua[n][r] * dq[n * TDIM + j];
    } else {
      for (int n = 0; n < nd; ++n) {
        const double* pu = a.u + (int64_t)ud[n] * TDIM;
#pragma unroll
        for (int r = 0; r < TDIM; ++r)
#pragma unroll
          for (int j = 0; j < TDIM; ++j) {
            const double t = pu[r] * dq[n * TDIM + j];
            H[r][j] = n == 0 ? t : H[r][j] + t;
          }
      }
    }
    double G[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int i = 0; i < 3; ++i) G[r][i] = 0.0;
#pragma unroll
    for (int r = 0; r < TDIM; ++r)
#pragma unroll
      for (int i = 0; i < TDIM; ++i) {
        double acc = H[r][0] * K[0][i];
#pragma unroll
        for (int j = 1; j < TDIM; ++j) acc = acc + H[r][j] * K[j][i];
        G[r][i] = acc;
      }
    const int64_t pt = c * a.nqp + q;
    double* o = a.out + pt;
    if (a.kind == 0) {
      o[0] = G[0][0];
      o[a.ld] = G[1][1];
      o[2 * a.ld] = G[2][2];
      o[3 * a.ld] = (G[0][1] + G[1][0]) * kR2;
      o[4 * a.ld] = (G[0][2] + G[2][0]) * kR2;
      o[5 * a.ld] = (G[1][2] + G[2][1]) * kR2;
    } else {
      o[0] = 1.0 + G[0][0];
      o[a.ld] = 1.0 + G[1][1];
      o[2 * a.ld] = 1.0 + G[2][2];
      o[3 * a.ld] = G[0][1];
      o[4 * a.ld] = G[1][0];
      o[5 * a.ld] = G[0][2];
      o[6 * a.ld] = G[2][0];
      o[7 * a.ld] = G[1][2];
      o[8 * a.ld] = G[2][1];
    }
  }
}

template <int TDIM, int ND>
__global__ void __launch_bounds__(128) fe_gradient_kernel(const FeGradArgs a) {
  __shared__ double s_dphi[kFeMaxTab];
  const int nd = ND > 0 ? ND : a.nd;
  const int ntab = a.nqp * nd * TDIM;
  for (int i = threadIdx.x; i < ntab && i < kFeMaxTab; i += blockDim.x) s_dphi[i] = a.dphi[i];
  __syncthreads();
  const double* dphi = ntab <= kFeMaxTab ? s_dphi : a.dphi;
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= a.num_cells) return;
  fe_gradient_cell<TDIM, ND>(a, dphi, nd, c);
}

}  // namespace dxm
