// CPU-side check of the FE kernels' per-cell / per-column routines (tests/test_fe_host.py): dxm::fe_gradient_cell,
// dxm::fe_form_point_geometry, fe_stage_flux / fe_stage_tangent, fe_form_column_g / fe_form_column_u, fe_form_entry and
// fe_form_vector_entry are __host__ __device__, so the very code the kernels run is
// executed here on the host and compared bit for bit with the oracles -- without a GPU.  The CTA-level staging of
// fe_forms_kernel (shared-memory arrays of the CTA's points) is replayed with one cell per "CTA".
// Test scaffolding only: nothing in the product calls this.
#include <cmath>
#include <cstring>
#include <vector>

#include "../dolfinx_materials_b200/csrc/dxm_fe_forms.cuh"
#include "../dolfinx_materials_b200/csrc/dxm_fe_gradient.cuh"

namespace {

long g_vector_mismatch = 0;  // fe_vector_node vs fe_form_vector_entry, entries that differ

template <int TDIM, int ND>
void grad_cells(const dxm::FeGradArgs& a) {
  const int nd = ND > 0 ? ND : a.nd;
  for (int64_t c = 0; c < a.num_cells; ++c) dxm::fe_gradient_cell<TDIM, ND>(a, a.dphi, nd, c);
}

// one cell per "CTA": the kernel's staging (tensor flux / tangent, g, vol g) and its per-lane column routines, column by column
template <int TDIM, int ND, int NQP>
void form_cells(const dxm::FeFormArgs& a) {
  constexpr int T2 = TDIM * TDIM;
  const int nd = ND > 0 ? ND : a.nd, ndof = nd * TDIM, nqp = a.nqp;
  const int nflux = a.kind == 0 ? 6 : 9, nct = a.kind == 0 ? dxm::kSym6Rows : 81;
  std::vector<double> vol(nqp), g((size_t)nqp * nd * TDIM), gv((size_t)nqp * nd * TDIM), S((size_t)nqp * T2),
      A((size_t)nqp * T2 * T2);
  double gb[dxm::kFeMaxQp * TDIM], U[dxm::kFeMaxQp * TDIM], Ut[dxm::kFeMaxQp * TDIM];
  for (int64_t c = 0; c < a.num_cells; ++c) {
    // K[m][j] row-major and |det J| of the cell, as the kernel keeps them in shared memory
    double Kc[TDIM][TDIM], detc, Kv[TDIM * TDIM + 1];
    dxm::cell_geometry<TDIM>(a.coords, a.geom_dofs + c * (TDIM + 1), Kc, detc);
    for (int i = 0; i < TDIM; ++i)
      for (int j = 0; j < TDIM; ++j) Kv[i * TDIM + j] = Kc[i][j];
    Kv[TDIM * TDIM] = std::fabs(detc);
    std::vector<double> gq((size_t)nd * TDIM);
    for (int q = 0; q < nqp; ++q) {
      dxm::fe_form_point_geometry<TDIM>(a, c, q, nd, vol[q], gq.data());
      for (int n = 0; n < nd; ++n)  // node-major staging, as the kernel lays g / vol g out in shared memory
        for (int j = 0; j < TDIM; ++j) {
          g[((size_t)n * nqp + q) * TDIM + j] = gq[n * TDIM + j];
          gv[((size_t)n * nqp + q) * TDIM + j] = vol[q] * gq[n * TDIM + j];
        }
      for (int row = 0; row < nflux; ++row)
        dxm::fe_stage_flux<TDIM>(a.kind, row, a.flux[(int64_t)row * a.ld + c * nqp + q], S.data() + (size_t)q * T2);
      if (a.want_mat)
        for (int row = 0; row < nct; ++row)
          dxm::fe_stage_tangent<TDIM>(a.kind, row, a.ct[(int64_t)row * a.ld + c * nqp + q], A.data() + (size_t)q * T2 * T2);
    }
    for (int col = 0; col < ndof; ++col) {
      const int b = col / TDIM, s = col % TDIM;
      dxm::fe_form_column_g<TDIM, NQP>(nqp, nd, b, g.data(), gb);
      a.fe[c * ndof + col] = dxm::fe_form_vector_entry<TDIM, NQP>(nqp, nd, b, s, gv.data(), S.data());
      if (a.want_mat)
        for (int r = 0; r < TDIM; ++r) {
          dxm::fe_form_column_u<TDIM, NQP>(nqp, r, s, gb, A.data(), U);
          dxm::fe_form_column_ut<TDIM, NQP>(nqp, Kv, a.weights, U, Ut);
          for (int an = 0; an < nd; ++an)
            a.ke[(c * ndof + an * TDIM + r) * ndof + col] = dxm::fe_form_entry<TDIM, NQP>(nqp, nd, an, a.dphi, Ut);
        }
    }
    // the residual-only kernel's per-(cell, basis function) routine must give the same bits as the column routine
    if (ND > 0 && NQP > 0)
      for (int n = 0; n < nd; ++n) {
        double acc[TDIM];
        dxm::fe_vector_node<TDIM, (ND > 0 ? ND : 1), (NQP > 0 ? NQP : 1)>(a, c, n, acc);
        for (int s = 0; s < TDIM; ++s)
          if (std::memcmp(&acc[s], &a.fe[c * ndof + n * TDIM + s], sizeof(double)) != 0) ++g_vector_mismatch;
      }
  }
}

}  // namespace

// generic != 0 forces the run-time-nd instantiation (ND = 0), otherwise the dispatch of dxm_fe_api.cu
extern "C" int fe_gradient_host(int tdim, int64_t num_cells, int nd, int nqp, int kind, const double* coords,
                                const int32_t* geom_dofs, const int32_t* u_dofs, const double* u, const double* dphi,
                                double* out, int64_t ld, int generic) {
  dxm::FeGradArgs a{coords, geom_dofs, u_dofs, u, dphi, out, ld, num_cells, nd, nqp, kind};
  if (tdim == 3) {
    if (!generic && nd == 4) return grad_cells<3, 4>(a), 0;
    if (!generic && nd == 10) return grad_cells<3, 10>(a), 0;
    return grad_cells<3, 0>(a), 0;
  }
  if (tdim == 2) {
    if (!generic && nd == 3) return grad_cells<2, 3>(a), 0;
    if (!generic && nd == 6) return grad_cells<2, 6>(a), 0;
    return grad_cells<2, 0>(a), 0;
  }
  return -1;
}

// flux [6|9][ld] and ct [21|81][ld] in the resident SoA layout (packed symmetric tangent for kind 0)
extern "C" int fe_forms_host(int tdim, int64_t num_cells, int nd, int nqp, int kind, const double* coords,
                             const int32_t* geom_dofs, const int32_t* u_dofs, const double* dphi, const double* weights,
                             const double* flux, const double* ct, int64_t ld, int want_mat, double* fe, double* ke,
                             int generic) {
  if (nd > dxm::kFeMaxNd) return -1;
  dxm::FeFormArgs a{};
  a.coords = coords;
  a.geom_dofs = geom_dofs;
  a.u_dofs = u_dofs;
  a.dphi = dphi;
  a.weights = weights;
  a.flux = flux;
  a.ct = ct;
  a.ld = ld;
  a.num_cells = num_cells;
  a.nd = nd;
  a.nqp = nqp;
  a.kind = kind;
  a.want_vec = 1;
  a.want_mat = want_mat;
  a.fe = fe;
  a.ke = ke;
  if (nqp > dxm::kFeMaxQp) return -1;
  g_vector_mismatch = 0;
  // same (nodes, Gauss points) dispatch as launch_fe_forms (dxm_fe_api.cu); -2: the residual-only routine disagrees
  if (tdim == 3) {
    if (!generic && nd == 10 && nqp == 4) form_cells<3, 10, 4>(a);
    else if (!generic && nd == 4 && nqp == 1) form_cells<3, 4, 1>(a);
    else if (!generic && nd == 4 && nqp == 4) form_cells<3, 4, 4>(a);
    else form_cells<3, 0, 0>(a);
    return g_vector_mismatch ? -2 : 0;
  }
  if (tdim == 2) {
    if (!generic && nd == 6 && nqp == 3) form_cells<2, 6, 3>(a);
    else if (!generic && nd == 3 && nqp == 1) form_cells<2, 3, 1>(a);
    else if (!generic && nd == 3 && nqp == 3) form_cells<2, 3, 3>(a);
    else form_cells<2, 0, 0>(a);
    return g_vector_mismatch ? -2 : 0;
  }
  return -1;
}
