// libdxm_cuda.so -- FE-side entry points: gradient evaluation, element forms, the
// device-resident assembled system and its Krylov stand-in.  Host side only; kernels live in the .cuh files.
#include "dxm_internal.cuh"
#include "dxm_fe_forms.cuh"
#include "dxm_fe_gradient.cuh"
#include "dxm_krylov.cuh"

using namespace dxm;

extern "C" {

// ---- FE gradient evaluation ---------------------------------------------------------
struct dxm_mesh {
  int device = 0, tdim = 3, nd = 4, nqp = 1;
  int64_t num_cells = 0, num_nodes = 0, num_dofs = 0;
  double *coords = nullptr, *dphi = nullptr, *u = nullptr, *weights = nullptr;
  int32_t *geom_dofs = nullptr, *u_dofs = nullptr;
  double *d_fe = nullptr, *d_ke = nullptr;  // element-form staging for host outputs (lazy)
  unsigned long long uid = 0;               // identity for caches keyed on a mesh (addresses get reused)
  std::vector<int32_t> h_u_dofs;  // host copy of the dofmap (node -> cells adjacency of the gather assembly)
  std::vector<double> h_dphi, h_weights;  // host copies: passed by value to fe_forms_kernel (constant-bank operands)
  int64_t* nc_ptr = nullptr;      // device: node -> cells around it (CSR), built on first use
  int32_t* nc_cell = nullptr;
  uint8_t* nc_loc = nullptr;
  bool adjacency = false;
};

int dxm_mesh_destroy(dxm_mesh* m) {
  if (!m) return 0;
  cudaSetDevice(m->device);
  cudaFree(m->coords);
  cudaFree(m->dphi);
  cudaFree(m->u);
  cudaFree(m->geom_dofs);
  cudaFree(m->u_dofs);
  cudaFree(m->weights);
  cudaFree(m->d_fe);
  cudaFree(m->d_ke);
  cudaFree(m->nc_ptr);
  cudaFree(m->nc_cell);
  cudaFree(m->nc_loc);
  delete m;
  return 0;
}

int dxm_mesh_create(int device, int tdim, int64_t num_cells, int64_t num_nodes, const double* coords,
                    const int32_t* geom_dofmap, int ndofs_cell, const int32_t* u_dofmap, int64_t num_dofs,
                    int nqp, const double* dphi, dxm_mesh** out) {
  if (!out) return fail("dxm_mesh_create: out is NULL");
  *out = nullptr;
  if (tdim != 2 && tdim != 3) return fail("dxm_mesh_create: tdim must be 2 or 3");
  if (num_cells <= 0 || num_nodes <= 0 || num_dofs <= 0 || ndofs_cell <= 0 || nqp <= 0 || !coords ||
      !geom_dofmap || !u_dofmap || !dphi)
    return fail("dxm_mesh_create: bad argument");
  CK(cudaSetDevice(device));
  static std::atomic<unsigned long long> next_uid{1};
  dxm_mesh* m = new dxm_mesh();
  m->uid = next_uid.fetch_add(1);
  m->device = device;
  m->tdim = tdim;
  m->nd = ndofs_cell;
  m->nqp = nqp;
  m->num_cells = num_cells;
  m->num_nodes = num_nodes;
  m->num_dofs = num_dofs;
  auto up = [&](void** d, const void* h, size_t bytes) -> cudaError_t {
    cudaError_t e = cudaMalloc(d, bytes);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice);
  };
  cudaError_t e = up((void**)&m->coords, coords, sizeof(double) * 3 * num_nodes);
  if (e == cudaSuccess) e = up((void**)&m->geom_dofs, geom_dofmap, sizeof(int32_t) * (tdim + 1) * num_cells);
  if (e == cudaSuccess) e = up((void**)&m->u_dofs, u_dofmap, sizeof(int32_t) * ndofs_cell * num_cells);
  m->h_u_dofs.assign(u_dofmap, u_dofmap + (size_t)ndofs_cell * num_cells);
  if (e == cudaSuccess) e = up((void**)&m->dphi, dphi, sizeof(double) * nqp * ndofs_cell * tdim);
  m->h_dphi.assign(dphi, dphi + (size_t)nqp * ndofs_cell * tdim);
  if (e == cudaSuccess) e = cudaMalloc((void**)&m->u, sizeof(double) * num_dofs * tdim);
  if (e != cudaSuccess) {
    dxm_mesh_destroy(m);
    return fail(std::string("dxm_mesh_create: ") + cudaGetErrorString(e));
  }
  *out = m;
  return 0;
}

int dxm_eval_gradient(dxm_mesh* m, dxm_handle* h, const double* u, int mem, int kind) {
  if (!m || !h || !u) return fail("dxm_eval_gradient: NULL argument");
  if (m->device != h->device) return fail("dxm_eval_gradient: mesh and material live on different devices");
  if (kind != 0 && kind != 1) return fail("dxm_eval_gradient: kind must be 0 (strain) or 1 (F)");
  if ((kind == 0 ? 6 : 9) != h->ngrad)
    return fail("dxm_eval_gradient: gradient kind does not match the behaviour's gradient size");
  if (m->num_cells * m->nqp != h->n)
    return fail("dxm_eval_gradient: num_cells*nqp = " + std::to_string(m->num_cells * m->nqp) +
                " but the material has " + std::to_string(h->n) + " Gauss points");
  if (mem != DXM_MEM_HOST && mem != DXM_MEM_DEVICE) return fail("dxm_eval_gradient: bad mem kind");
  if (set_device(h)) return -1;
  CK(cudaMemcpyAsync(m->u, u, sizeof(double) * m->num_dofs * m->tdim,
                     mem == DXM_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, h->stream));
  FeGradArgs a{};
  a.coords = m->coords;
  a.geom_dofs = m->geom_dofs;
  a.u_dofs = m->u_dofs;
  a.u = m->u;
  a.dphi = m->dphi;
  a.out = h->gen[1 - h->i0];  // s1's gradient block
  a.ld = h->ld;
  a.num_cells = m->num_cells;
  a.nd = m->nd;
  a.nqp = m->nqp;
  a.kind = kind;
  const int block = 128;
  const int grid = (int)((m->num_cells + block - 1) / block);
  if (m->tdim == 3) {
    if (m->nd == 4)
      fe_gradient_kernel<3, 4><<<grid, block, 0, h->stream>>>(a);
    else if (m->nd == 10)
      fe_gradient_kernel<3, 10><<<grid, block, 0, h->stream>>>(a);
    else
      fe_gradient_kernel<3, 0><<<grid, block, 0, h->stream>>>(a);
  } else {
    if (m->nd == 3)
      fe_gradient_kernel<2, 3><<<grid, block, 0, h->stream>>>(a);
    else if (m->nd == 6)
      fe_gradient_kernel<2, 6><<<grid, block, 0, h->stream>>>(a);
    else
      fe_gradient_kernel<2, 0><<<grid, block, 0, h->stream>>>(a);
  }
  LAUNCH_CHECK();
  if (mem == DXM_MEM_HOST) CK(cudaStreamSynchronize(h->stream));  // the caller's u is free again
  return 0;
}

// ---- fused flux / tangent -> element forms and global assembly -----------------------
struct dxm_system {
  int device = 0;
  int64_t nrows = 0, nnz = 0;
  int64_t* rowptr = nullptr;
  int32_t* colidx = nullptr;
  double *vals = nullptr, *rhs = nullptr;
  uint8_t* bc = nullptr;
  double* lift = nullptr;  // prescribed solution values on the constrained dofs
  bool lift_on = false;    // false: homogeneous
  bool last_lift_on = false, last_vec = false, last_mat = false;  // what the last dxm_assemble produced
  bool defer_bc = false;   // true: dxm_assemble leaves the constrained rows to dxm_system_apply_constraints
  unsigned long long* missing = nullptr;
  double* work = nullptr;  // Krylov workspace: 9 vectors + block inverses + scalar slots (lazy)
  // per-cell block offsets into the pattern, built at the first assembly with a given mesh (node-blocked patterns)
  int32_t* off = nullptr;
  unsigned long long off_mesh = 0;  // uid of the mesh the table was built for
  bool off_valid = false;
  int max_row_len = 0;  // longest CSR row (sizes the per-node row image of the gather assembly)
};

int dxm_mesh_set_weights(dxm_mesh* m, const double* weights) {
  if (!m || !weights) return fail("dxm_mesh_set_weights: NULL argument");
  CK(cudaSetDevice(m->device));
  if (!m->weights) CK(cudaMalloc((void**)&m->weights, sizeof(double) * m->nqp));
  CK(cudaMemcpy(m->weights, weights, sizeof(double) * m->nqp, cudaMemcpyHostToDevice));
  m->h_weights.assign(weights, weights + m->nqp);
  return 0;
}

extern "C++" {
namespace {

int check_forms(const char* who, dxm_mesh* m, dxm_handle* h, int kind) {
  if (!m || !h) return fail(std::string(who) + ": NULL argument");
  if (m->device != h->device) return fail(std::string(who) + ": mesh and material live on different devices");
  if (kind != 0 && kind != 1) return fail(std::string(who) + ": kind must be 0 (strain/stress) or 1 (F/PK1)");
  if ((kind == 0 ? 6 : 9) != h->ngrad)
    return fail(std::string(who) + ": kind does not match the behaviour's gradient size");
  if (m->num_cells * m->nqp != h->n)
    return fail(std::string(who) + ": num_cells*nqp = " + std::to_string(m->num_cells * m->nqp) +
                " but the material has " + std::to_string(h->n) + " Gauss points");
  if (!m->weights) return fail(std::string(who) + ": quadrature weights not set (dxm_mesh_set_weights)");
  if (m->nd > kFeMaxNd) return fail(std::string(who) + ": at most " + std::to_string(kFeMaxNd) + " dofs per cell");
  if (m->nqp > kFeMaxQp) return fail(std::string(who) + ": at most " + std::to_string(kFeMaxQp) + " Gauss points per cell");
  if (!h->s1_valid && h->last.n_points == 0)
    return fail(std::string(who) + ": no constitutive update has been run on this material yet");
  return 0;
}

// node -> (cell, local index) adjacency of the mesh, cells in ascending order (the fixed summation order of the gather)
int build_adjacency(dxm_mesh* m) {
  if (m->adjacency) return 0;
  const int64_t nc = m->num_cells, nn = m->num_dofs;
  const int nd = m->nd;
  std::vector<int64_t> ptr((size_t)nn + 1, 0);
  for (int64_t i = 0; i < nc * nd; ++i) ++ptr[(size_t)m->h_u_dofs[i] + 1];
  for (int64_t n = 0; n < nn; ++n) ptr[n + 1] += ptr[n];
  std::vector<int32_t> cell((size_t)nc * nd);
  std::vector<uint8_t> loc((size_t)nc * nd);
  std::vector<int64_t> next(ptr.begin(), ptr.end() - 1);
  for (int64_t c = 0; c < nc; ++c)
    for (int a = 0; a < nd; ++a) {
      const int64_t k = next[m->h_u_dofs[c * nd + a]]++;
      cell[k] = (int32_t)c;
      loc[k] = (uint8_t)a;
    }
  CK(cudaMalloc((void**)&m->nc_ptr, sizeof(int64_t) * (nn + 1)));
  CK(cudaMalloc((void**)&m->nc_cell, sizeof(int32_t) * nc * nd));
  CK(cudaMalloc((void**)&m->nc_loc, (size_t)nc * nd));
  CK(cudaMemcpy(m->nc_ptr, ptr.data(), sizeof(int64_t) * (nn + 1), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(m->nc_cell, cell.data(), sizeof(int32_t) * nc * nd, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(m->nc_loc, loc.data(), (size_t)nc * nd, cudaMemcpyHostToDevice));
  m->adjacency = true;
  return 0;
}

// one instantiation: the reference gradients / weights of the element travel as kernel parameters (by value)
template <int TDIM, int ND, int NQP, int MODE>
int launch_fe_forms_inst(dxm_mesh* m, dxm_handle* h, const FeFormArgs& a, const FeFormSmem& L, int64_t grid) {
  FeTab<ND * NQP * TDIM, NQP> T{};
  if (ND > 0) {
    for (int i = 0; i < ND * NQP * TDIM; ++i) T.v[i] = m->h_dphi[i];
    for (int q = 0; q < NQP; ++q) T.w[q] = m->h_weights[q];
  }
  if (L.bytes > 48 * 1024)
    CK(cudaFuncSetAttribute(fe_forms_kernel<TDIM, ND, NQP, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)L.bytes));
  fe_forms_kernel<TDIM, ND, NQP, MODE><<<(unsigned)grid, 32 * kFeWarps, L.bytes, h->stream>>>(a, L, T);
  LAUNCH_CHECK();
  return 0;
}

template <int MODE>
int launch_fe_forms(dxm_mesh* m, dxm_handle* h, FeFormArgs& a) {
  if (!a.want_mat && a.want_vec) {
    // residual only: the thread-per-(cell, basis function) kernel where (nodes, Gauss points) are compile-time
    const void* vk = nullptr;
    if (m->tdim == 3)
      vk = (m->nd == 10 && m->nqp == 4)  ? (const void*)fe_vector_kernel<3, 10, 4, MODE>
           : (m->nd == 4 && m->nqp == 1) ? (const void*)fe_vector_kernel<3, 4, 1, MODE>
           : (m->nd == 4 && m->nqp == 4) ? (const void*)fe_vector_kernel<3, 4, 4, MODE>
                                         : nullptr;
    else
      vk = (m->nd == 6 && m->nqp == 3)   ? (const void*)fe_vector_kernel<2, 6, 3, MODE>
           : (m->nd == 3 && m->nqp == 1) ? (const void*)fe_vector_kernel<2, 3, 1, MODE>
           : (m->nd == 3 && m->nqp == 3) ? (const void*)fe_vector_kernel<2, 3, 3, MODE>
                                         : nullptr;
    if (vk) {
      const int64_t grid = (a.num_cells * m->nd + kFeVecBlock - 1) / kFeVecBlock;
      if (grid > 0x7fffffff) return fail("fe_forms: too many cells for one launch");
      void* vargs[] = {(void*)&a};
      CK(cudaLaunchKernel(vk, dim3((unsigned)grid), dim3(kFeVecBlock), vargs, 0, h->stream));
      LAUNCH_CHECK();
      return 0;
    }
  }
  const FeFormSmem L = fe_form_smem(m->tdim, m->nd, m->nqp, a.kind, MODE, a.want_mat != 0);
  if (L.bytes > 227 * 1024) return fail("fe_forms: element too large for the shared-memory staging");
  const int64_t grid = (a.num_cells + L.cpb - 1) / L.cpb;
  if (grid > 0x7fffffff) return fail("fe_forms: too many cells for one launch");
  // compile-time (nodes, Gauss points) for the hot-path elements: P2 / degree 2 and P1 / degree <= 1 simplices; anything
  // else runs the run-time instantiation
  if (m->tdim == 3) {
    if (m->nd == 10 && m->nqp == 4) return launch_fe_forms_inst<3, 10, 4, MODE>(m, h, a, L, grid);
    if (m->nd == 4 && m->nqp == 1) return launch_fe_forms_inst<3, 4, 1, MODE>(m, h, a, L, grid);
    if (m->nd == 4 && m->nqp == 4) return launch_fe_forms_inst<3, 4, 4, MODE>(m, h, a, L, grid);
    return launch_fe_forms_inst<3, 0, 0, MODE>(m, h, a, L, grid);
  }
  if (m->nd == 6 && m->nqp == 3) return launch_fe_forms_inst<2, 6, 3, MODE>(m, h, a, L, grid);
  if (m->nd == 3 && m->nqp == 1) return launch_fe_forms_inst<2, 3, 1, MODE>(m, h, a, L, grid);
  if (m->nd == 3 && m->nqp == 3) return launch_fe_forms_inst<2, 3, 3, MODE>(m, h, a, L, grid);
  return launch_fe_forms_inst<2, 0, 0, MODE>(m, h, a, L, grid);
}

void fill_form_args(dxm_mesh* m, dxm_handle* h, int kind, FeFormArgs& a) {
  // flux and tangent of the last update: s1 while it is valid, else the generation it became at update()
  const int g = h->s1_valid ? 1 - h->i0 : h->i0;
  a.coords = m->coords;
  a.geom_dofs = m->geom_dofs;
  a.u_dofs = m->u_dofs;
  a.dphi = m->dphi;
  a.weights = m->weights;
  a.flux = h->gen[g] + (int64_t)h->ngrad * h->ld;
  a.ct = h->ct;
  a.ld = h->ld;
  a.num_cells = m->num_cells;
  a.nd = m->nd;
  a.nqp = m->nqp;
  a.kind = kind;
}

}  // namespace
}  // extern "C++"

int dxm_element_forms(dxm_mesh* m, dxm_handle* h, int kind, double* fe, double* ke, int mem) {
  if (check_forms("dxm_element_forms", m, h, kind)) return -1;
  if (mem != DXM_MEM_HOST && mem != DXM_MEM_DEVICE) return fail("dxm_element_forms: bad mem kind");
  if (!fe && !ke) return 0;
  if (set_device(h)) return -1;
  const int64_t ndof = (int64_t)m->nd * m->tdim;
  const size_t fe_bytes = sizeof(double) * m->num_cells * ndof, ke_bytes = fe_bytes * ndof;
  FeFormArgs a{};
  fill_form_args(m, h, kind, a);
  a.want_vec = fe != nullptr;
  a.want_mat = ke != nullptr;
  if (mem == DXM_MEM_DEVICE) {
    a.fe = fe;
    a.ke = ke;
  } else {
    if (fe && !m->d_fe) CK(cudaMalloc((void**)&m->d_fe, fe_bytes));
    if (ke && !m->d_ke) CK(cudaMalloc((void**)&m->d_ke, ke_bytes));
    a.fe = m->d_fe;
    a.ke = m->d_ke;
  }
  if (launch_fe_forms<MODE_ELEMENT>(m, h, a)) return -1;
  if (mem == DXM_MEM_HOST) {
    if (fe) CK(cudaMemcpyAsync(fe, m->d_fe, fe_bytes, cudaMemcpyDeviceToHost, h->stream));
    if (ke) CK(cudaMemcpyAsync(ke, m->d_ke, ke_bytes, cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int dxm_system_destroy(dxm_system* s) {
  if (!s) return 0;
  cudaSetDevice(s->device);
  cudaFree(s->rowptr);
  cudaFree(s->colidx);
  cudaFree(s->vals);
  cudaFree(s->rhs);
  cudaFree(s->bc);
  cudaFree(s->missing);
  cudaFree(s->lift);
  cudaFree(s->work);
  cudaFree(s->off);
  delete s;
  return 0;
}

int dxm_system_create(int device, int64_t nrows, const int64_t* rowptr, const int32_t* colidx, dxm_system** out) {
  if (!out) return fail("dxm_system_create: out is NULL");
  *out = nullptr;
  if (nrows <= 0 || !rowptr || !colidx) return fail("dxm_system_create: bad argument");
  if (rowptr[0] != 0 || rowptr[nrows] <= 0) return fail("dxm_system_create: rowptr must start at 0 and be non-empty");
  CK(cudaSetDevice(device));
  dxm_system* s = new dxm_system();
  s->device = device;
  s->nrows = nrows;
  s->nnz = rowptr[nrows];
  for (int64_t i = 0; i < nrows; ++i) s->max_row_len = std::max<int64_t>(s->max_row_len, rowptr[i + 1] - rowptr[i]);
  cudaError_t e = cudaMalloc((void**)&s->rowptr, sizeof(int64_t) * (nrows + 1));
  if (e == cudaSuccess) e = cudaMemcpy(s->rowptr, rowptr, sizeof(int64_t) * (nrows + 1), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->colidx, sizeof(int32_t) * s->nnz);
  if (e == cudaSuccess) e = cudaMemcpy(s->colidx, colidx, sizeof(int32_t) * s->nnz, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->vals, sizeof(double) * s->nnz);
  if (e == cudaSuccess) e = cudaMemset(s->vals, 0, sizeof(double) * s->nnz);
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->rhs, sizeof(double) * nrows);
  if (e == cudaSuccess) e = cudaMemset(s->rhs, 0, sizeof(double) * nrows);
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->missing, sizeof(unsigned long long));
  if (e != cudaSuccess) {
    dxm_system_destroy(s);
    return fail(std::string("dxm_system_create: ") + cudaGetErrorString(e));
  }
  *out = s;
  return 0;
}

int64_t dxm_system_nnz(const dxm_system* s) { return s ? s->nnz : -1; }

int dxm_system_set_bc(dxm_system* s, const uint8_t* marker) {
  if (!s) return fail("dxm_system_set_bc: NULL system");
  CK(cudaSetDevice(s->device));
  if (!marker) {
    cudaFree(s->bc);
    s->bc = nullptr;
    return 0;
  }
  if (!s->bc) CK(cudaMalloc((void**)&s->bc, s->nrows));
  CK(cudaMemcpy(s->bc, marker, s->nrows, cudaMemcpyHostToDevice));
  return 0;
}

int dxm_system_set_lifting(dxm_system* s, const double* values) {
  if (!s) return fail("dxm_system_set_lifting: NULL system");
  CK(cudaSetDevice(s->device));
  if (!values) {
    s->lift_on = false;
    return 0;
  }
  if (!s->lift) CK(cudaMalloc((void**)&s->lift, sizeof(double) * s->nrows));
  CK(cudaMemcpy(s->lift, values, sizeof(double) * s->nrows, cudaMemcpyHostToDevice));
  s->lift_on = true;
  return 0;
}

int dxm_assemble(dxm_mesh* m, dxm_handle* h, int kind, dxm_system* s, int want_vector, int want_matrix) {
  if (check_forms("dxm_assemble", m, h, kind)) return -1;
  if (!s) return fail("dxm_assemble: NULL system");
  if (s->device != h->device) return fail("dxm_assemble: system and material live on different devices");
  if (s->nrows != m->num_dofs * m->tdim)
    return fail("dxm_assemble: the system has " + std::to_string(s->nrows) + " rows but the space has " +
                std::to_string(m->num_dofs * m->tdim) + " dofs");
  if (!want_vector && !want_matrix) return 0;
  if (s->lift_on && s->bc && want_vector && !want_matrix)
    return fail("dxm_assemble: lifting the constrained columns needs the matrix pass (want_matrix)");
  if (set_device(h)) return -1;
  FeFormArgs a{};
  fill_form_args(m, h, kind, a);
  a.want_vec = want_vector != 0;
  a.want_mat = want_matrix != 0;
  a.b = s->rhs;
  a.rowptr = s->rowptr;
  a.colidx = s->colidx;
  a.vals = s->vals;
  a.bc = s->bc;
  a.lift = (s->bc && s->lift_on && want_vector && want_matrix) ? s->lift : nullptr;
  a.missing = s->missing;
  if (want_matrix && s->off_mesh != m->uid) {
    // first assembly of this mesh into this pattern: locate every (cell, node a, node b) block once
    cudaFree(s->off);
    s->off = nullptr;
    s->off_valid = false;
    s->off_mesh = m->uid;
    const int64_t cnt = m->num_cells * m->nd * m->nd;
    if (cudaMalloc((void**)&s->off, sizeof(int32_t) * cnt) == cudaSuccess) {
      CK(cudaMemsetAsync(s->missing, 0, sizeof(unsigned long long), h->stream));
      fe_offsets_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, h->stream>>>(m->u_dofs, m->num_cells, m->nd, m->tdim,
                                                                            s->rowptr, s->colidx, s->off, s->missing);
      LAUNCH_CHECK();
      unsigned long long bad = 0;
      CK(cudaMemcpyAsync(&bad, s->missing, sizeof(bad), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      s->off_valid = bad == 0;  // otherwise: not node-blocked or entries missing -> per-entry search path reports
    } else {
      cudaGetLastError();  // no room for the table: keep searching
    }
  }
  a.off = (want_matrix && s->off_valid) ? s->off : nullptr;
  // Default: the fused contraction + atomic scatter (6.5 ms for 663 k P2 tetrahedra, bound by the ~90 G fp64
  // reductions / s the L2 retires).  DXM_FE_GATHER=1: element matrices as full lines + the per-node gather -- no atomics,
  // fixed summation order (bit-reproducible assembly), whole CSR rows written once -- at 7.8 ms with 4.9 GB of scratch;
  // needs the block-offset table, the element-matrix buffer and a row image that fits shared memory.
  const char* ge = std::getenv("DXM_FE_GATHER");
  const bool gather_on = ge && std::atoi(ge) != 0;
  const int64_t ndof = (int64_t)m->nd * m->tdim;
  const size_t gather_smem = sizeof(double) * kGatherWarps * m->tdim * (size_t)s->max_row_len;
  bool gather = want_matrix && gather_on && a.off && m->num_cells < 0x7fffffff && m->nd < 256 && gather_smem <= 200 * 1024;
  if (gather && build_adjacency(m)) return -1;
  if (gather && !m->d_ke && cudaMalloc((void**)&m->d_ke, sizeof(double) * m->num_cells * ndof * ndof) != cudaSuccess) {
    cudaGetLastError();
    gather = false;  // no room for the element matrices
  }
  if (gather && !m->d_fe) CK(cudaMalloc((void**)&m->d_fe, sizeof(double) * m->num_cells * ndof));
  CK(cudaMemsetAsync(s->missing, 0, sizeof(unsigned long long), h->stream));
  if (gather) {
    FeFormArgs e = a;
    e.want_vec = 1;
    e.fe = m->d_fe;
    e.ke = m->d_ke;
    if (launch_fe_forms<MODE_ELEMENT>(m, h, e)) return -1;
    FeGatherArgs g{};
    g.ke = m->d_ke;
    g.fe = m->d_fe;
    g.nc_ptr = m->nc_ptr;
    g.nc_cell = m->nc_cell;
    g.nc_loc = m->nc_loc;
    g.off = s->off;
    g.rowptr = s->rowptr;
    g.colidx = s->colidx;
    g.vals = s->vals;
    g.rhs = s->rhs;
    g.bc = s->bc;
    g.lift = a.lift;
    g.num_nodes = m->num_dofs;
    g.nd = m->nd;
    g.want_vec = want_vector != 0;
    g.maxlen = s->max_row_len;
    const void* gk = m->tdim == 3 ? (const void*)fe_gather_kernel<3> : (const void*)fe_gather_kernel<2>;
    if (gather_smem > 48 * 1024) CK(cudaFuncSetAttribute(gk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gather_smem));
    void* gargs[] = {(void*)&g};
    CK(cudaLaunchKernel(gk, dim3((unsigned)((m->num_dofs + kGatherWarps - 1) / kGatherWarps)), dim3(32 * kGatherWarps), gargs,
                        gather_smem, h->stream));
    LAUNCH_CHECK();
  } else {
    if (want_vector) CK(cudaMemsetAsync(s->rhs, 0, sizeof(double) * s->nrows, h->stream));
    if (want_matrix) CK(cudaMemsetAsync(s->vals, 0, sizeof(double) * s->nnz, h->stream));
    if (launch_fe_forms<MODE_GLOBAL>(m, h, a)) return -1;
  }
  s->last_lift_on = a.lift != nullptr;
  s->last_vec = want_vector != 0;
  s->last_mat = want_matrix != 0;
  if (s->bc && !s->defer_bc) {
    fe_bc_diag_kernel<<<(unsigned)((s->nrows + 255) / 256), 256, 0, h->stream>>>(
        s->bc, s->rowptr, s->colidx, want_matrix ? s->vals : nullptr, s->nrows, a.lift,
        want_vector ? s->rhs : nullptr);
    LAUNCH_CHECK();
  }
  unsigned long long miss = 0;
  CK(cudaMemcpyAsync(&miss, s->missing, sizeof(miss), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (miss)
    return fail("dxm_assemble: " + std::to_string(miss) + " element entries have no slot in the CSR pattern");
  return 0;
}

int dxm_system_defer_constraints(dxm_system* s, int on) {
  if (!s) return fail("dxm_system_defer_constraints: NULL system");
  s->defer_bc = on != 0;
  return 0;
}

int dxm_system_apply_constraints(dxm_system* s) {
  if (!s) return fail("dxm_system_apply_constraints: NULL system");
  if (!s->bc) return 0;
  CK(cudaSetDevice(s->device));
  fe_bc_diag_kernel<<<(unsigned)((s->nrows + 255) / 256), 256>>>(s->bc, s->rowptr, s->colidx,
                                                               s->last_mat ? s->vals : nullptr, s->nrows,
                                                               s->last_lift_on ? s->lift : nullptr,
                                                               s->last_vec ? s->rhs : nullptr);
  LAUNCH_CHECK();
  CK(cudaDeviceSynchronize());
  return 0;
}

int dxm_system_device_ptrs(dxm_system* s, double** values, double** rhs) {
  if (!s) return fail("dxm_system_device_ptrs: NULL system");
  if (values) *values = s->vals;
  if (rhs) *rhs = s->rhs;
  return 0;
}

int dxm_system_get(dxm_system* s, double* values, double* rhs, int mem) {
  if (!s) return fail("dxm_system_get: NULL system");
  if (mem != DXM_MEM_HOST && mem != DXM_MEM_DEVICE) return fail("dxm_system_get: bad mem kind");
  CK(cudaSetDevice(s->device));
  const cudaMemcpyKind kd = mem == DXM_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  if (values) CK(cudaMemcpy(values, s->vals, sizeof(double) * s->nnz, kd));
  if (rhs) CK(cudaMemcpy(rhs, s->rhs, sizeof(double) * s->nrows, kd));
  return 0;
}

int dxm_system_solve(dxm_system* s, int block, double rtol, int maxit, double* x, int mem, int* iters,
                     double* relres) {
  if (!s || !x) return fail("dxm_system_solve: NULL argument");
  if (block < 1 || block > 3 || s->nrows % block) return fail("dxm_system_solve: block must be 1, 2 or 3 and divide nrows");
  if (mem != DXM_MEM_HOST && mem != DXM_MEM_DEVICE) return fail("dxm_system_solve: bad mem kind");
  if (maxit < 1) return fail("dxm_system_solve: maxit must be positive");
  CK(cudaSetDevice(s->device));
  const int64_t n = s->nrows;
  const int64_t npad = (n + 31) & ~int64_t(31);
  if (!s->work) CK(cudaMalloc((void**)&s->work, sizeof(double) * (npad * 9 + n * 3 + 2 * KS_N + 4)));
  KrylovVecs k{};
  double* w = s->work;
  k.x = w;
  k.r = w + npad;
  k.rhat = w + 2 * npad;
  k.p = w + 3 * npad;
  k.v = w + 4 * npad;
  k.s = w + 5 * npad;
  k.t = w + 6 * npad;
  k.y = w + 7 * npad;
  k.z = w + 8 * npad;
  double* minv = w + 9 * npad;
  k.minv = minv;
  k.slots = minv + n * 3;
  k.scal = k.slots + 2 * KS_N;
  k.n = n;
  k.bs = block;
  cudaStream_t st = nullptr;  // legacy default stream: ordered after the assembly's stream synchronisation
  const unsigned gv = (unsigned)((n + 255) / 256);
  block_jacobi_kernel<<<(unsigned)((n / block + 255) / 256), 256, 0, st>>>(s->rowptr, s->colidx, s->vals, n / block,
                                                                          block, minv);
  LAUNCH_CHECK();
  CK(cudaMemsetAsync(k.slots, 0, sizeof(double) * 2 * KS_N, st));
  const double one[4] = {1.0, 1.0, 1.0, 0.0};
  CK(cudaMemcpyAsync(k.scal, one, sizeof(one), cudaMemcpyHostToDevice, st));
  bicg_init_kernel<<<gv, 256, 0, st>>>(k, s->rhs);
  LAUNCH_CHECK();
  double bb = 0.0;
  CK(cudaMemcpy(&bb, k.slots + KS_RR, sizeof(double), cudaMemcpyDeviceToHost));
  int it = 0;
  double rel = 0.0;
  if (bb > 0.0 && std::isfinite(bb)) {
    const double avg = (double)s->nnz / (double)n;
    const int G = avg < 24 ? 4 : avg < 48 ? 8 : avg < 96 ? 16 : 32;
    const unsigned gs = (unsigned)((n * G + 255) / 256);
    auto spmv = [&](const double* in, double* out, const double* d1, double* s1, const double* d2, double* s2) {
      switch (G) {
        case 4: spmv_dot_kernel<4><<<gs, 256, 0, st>>>(s->rowptr, s->colidx, s->vals, in, out, n, d1, s1, d2, s2); break;
        case 8: spmv_dot_kernel<8><<<gs, 256, 0, st>>>(s->rowptr, s->colidx, s->vals, in, out, n, d1, s1, d2, s2); break;
        case 16: spmv_dot_kernel<16><<<gs, 256, 0, st>>>(s->rowptr, s->colidx, s->vals, in, out, n, d1, s1, d2, s2); break;
        default: spmv_dot_kernel<32><<<gs, 256, 0, st>>>(s->rowptr, s->colidx, s->vals, in, out, n, d1, s1, d2, s2); break;
      }
      g_launches.fetch_add(1);
    };
    const int check_every = 8;
    rel = 1.0;
    while (it < maxit) {
      const int bank = it & 1;
      double* S = k.slots + bank * KS_N;
      bicg_p_kernel<<<gv, 256, 0, st>>>(k, bank);
      bicg_prec_kernel<<<gv, 256, 0, st>>>(k, k.p, k.y);
      spmv(k.y, k.v, k.rhat, S + KS_RV, nullptr, nullptr);
      bicg_s_kernel<<<gv, 256, 0, st>>>(k, bank);
      bicg_prec_kernel<<<gv, 256, 0, st>>>(k, k.s, k.z);
      spmv(k.z, k.t, k.s, S + KS_TS, nullptr, S + KS_TT);
      bicg_x_kernel<<<gv, 256, 0, st>>>(k, bank);
      bicg_scal_kernel<<<1, 1, 0, st>>>(k, bank);
      g_launches.fetch_add(6);
      ++it;
      if (it % check_every == 0 || it == maxit) {
        double rr = 0.0;
        CK(cudaMemcpy(&rr, k.slots + (1 - bank) * KS_N + KS_RR, sizeof(double), cudaMemcpyDeviceToHost));
        rel = std::sqrt(rr / bb);
        if (!std::isfinite(rel)) break;
        if (rel <= rtol) break;
      }
    }
    CK(cudaGetLastError());
  }
  CK(cudaMemcpy(x, k.x, sizeof(double) * n, mem == DXM_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice));
  if (iters) *iters = it;
  if (relres) *relres = rel;
  if (!std::isfinite(rel)) return fail("dxm_system_solve: BiCGStab broke down (non-finite residual)");
  return rel <= rtol ? 0 : 1;
}

}  // extern "C"
