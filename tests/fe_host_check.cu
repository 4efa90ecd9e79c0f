// CPU-side check of the FE kernels' per-cell / per-row routines (tests/test_fe_host.py): dxm::fe_gradient_cell,
// dxm::fe_form_point_geometry and dxm::fe_form_row are __host__ __device__, so the very code the kernels run is
// executed here on the host and compared bit for bit with the oracles -- without a GPU.  The CTA-level staging of
// fe_forms_kernel (shared-memory arrays of the CTA's points) is replayed with one cell per "CTA".
// Test scaffolding only: nothing in the product calls this.
#include <cmath>
#include <vector>

#include "../dolfinx_materials_b200/csrc/dxm_fe_forms.cuh"
#include "../dolfinx_materials_b200/csrc/dxm_fe_gradient.cuh"

namespace {

template <int TDIM, int ND>
void grad_cells(const dxm::FeGradArgs& a) {
  const int nd = ND > 0 ? ND : a.nd;
  for (int64_t c = 0; c < a.num_cells; ++c) dxm::fe_gradient_cell<TDIM, ND>(a, a.dphi, nd, c);
}

template <int TDIM, int ND>
void form_cells(const dxm::FeFormArgs& a) {
  const int nd = ND > 0 ? ND : a.nd, ndof = nd * TDIM, nqp = a.nqp, np = nqp;
  const int nflux = a.kind == 0 ? 6 : 9, nct = a.kind == 0 ? dxm::kSym6Rows : 81;
  std::vector<double> vol(np), g((size_t)np * nd * TDIM), fl((size_t)nflux * np), ct((size_t)nct * np);
  std::vector<double> acc(dxm::kFeMaxNd * TDIM);
  for (int64_t c = 0; c < a.num_cells; ++c) {
    for (int q = 0; q < nqp; ++q) dxm::fe_form_point_geometry<TDIM>(a, c, q, nd, vol[q], g.data() + (size_t)q * nd * TDIM);
    for (int row = 0; row < nflux; ++row)
      for (int q = 0; q < nqp; ++q) fl[(size_t)row * np + q] = a.flux[(int64_t)row * a.ld + c * nqp + q];
    if (a.want_mat)
      for (int row = 0; row < nct; ++row)
        for (int q = 0; q < nqp; ++q) ct[(size_t)row * np + q] = a.ct[(int64_t)row * a.ld + c * nqp + q];
    for (int row = 0; row < ndof; ++row) {
      double fe = 0.0;
      dxm::fe_form_row<TDIM, ND>(a.kind, a.want_mat != 0, nqp, nd, np, 0, row / TDIM, row % TDIM, vol.data(), g.data(),
                                 fl.data(), ct.data(), fe, acc.data());
      a.fe[c * ndof + row] = fe;
      if (a.want_mat)
        for (int k = 0; k < ndof; ++k) a.ke[(c * ndof + row) * ndof + k] = acc[k];
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
  if (tdim == 3) {
    if (!generic && nd == 4) return form_cells<3, 4>(a), 0;
    if (!generic && nd == 10) return form_cells<3, 10>(a), 0;
    return form_cells<3, 0>(a), 0;
  }
  if (tdim == 2) {
    if (!generic && nd == 3) return form_cells<2, 3>(a), 0;
    if (!generic && nd == 6) return form_cells<2, 6>(a), 0;
    return form_cells<2, 0>(a), 0;
  }
  return -1;
}
