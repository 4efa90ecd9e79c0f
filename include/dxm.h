/*
 * dxm.h -- C ABI of libdxm_cuda.so: the B200-native batched constitutive update behind the
 * dolfinx_materials `Material` protocol.
 *
 * The reference (bleyerj/dolfinx_materials v0.4.0) is pure Python and has no FFI of its own for
 * this path; each entry point below names the reference interface it replaces
 * (paths relative to the upstream tree).  The closest native precedent in the reference is the
 * MGIS call `mgis_bv.integrate(data_manager, type, dt, begin, end) -> int status`
 * (dolfinx_materials/mfront.py:266-272), whose "status < 1 => warning" convention is kept.
 *
 * Conventions
 *  - every function returns 0 on success, < 0 on a hard error (text via dxm_last_error()),
 *    dxm_integrate additionally returns > 0 = number of Gauss points whose local solve failed
 *    (iteration cap or non-finite result; replaces the host NaN scans of quadrature_map.py:322-324).
 *  - nothing throws or aborts across the ABI; plain pointers and sizes only.
 *  - the library owns all device buffers of a handle (two state generations, tangent, staging);
 *    the caller owns every array it passes, borrowed for the duration of the call.
 *  - host/“AoS” arrays are C-contiguous (n, dim) float64 exactly as QuadratureMap builds and consumes
 *    them (quadrature_map.py:313, :331-334; utils.py:136-143).  Device-resident fields are SoA:
 *    component c of point i at  base[c * ld + i],  ld = dxm_ld(h).  The resident tangent of the small-strain
 *    behaviours is symmetric and stored packed: 21 rows, entry (j, i), j <= i, of the row-major 6x6 at row
 *    j*6 - j(j-1)/2 + (i-j); every host / device-AoS output is the full (n, 36) array (mirrored on the way out).
 *    The finite-strain tangent is stored in full (81 rows).
 *  - tensor conventions: symmetric tensors are Mandel 6-vectors [11,22,33,r2*12,r2*13,r2*23]
 *    (utils.py:146-165), non-symmetric ones [11,22,33,12,21,13,31,23,32] (utils.py:168-190);
 *    the tangent is row-major d flux_j / d grad_i at j*ngrad+i (quadrature_map.py:94-104).
 */
#ifndef DXM_H
#define DXM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dxm_handle dxm_handle;

/* behaviours (reference: the jaxmat behaviours wrapped by JAXMaterial, jaxmat.py:141-234, and
 * python_materials/elasticity.py:5-24) */
enum {
  DXM_ELASTIC = 0,   /* LinearElasticIsotropic: E, nu                                              */
  DXM_J2_LINEAR = 1, /* J2 + linear isotropic hardening, closed form: E, nu, sig0, H                */
  DXM_J2_VOCE = 2,   /* J2 + sig0 + H p + (sigu-sig0)(1-exp(-b p)), scalar Newton: E,nu,sig0,sigu,b,H */
  DXM_FEFP_VOCE = 3, /* finite-strain FeFp J2 plasticity, same hardening law                        */
  DXM_J2_TABLE = 4,  /* J2 + piecewise-linear isotropic hardening table (dxm_set_hardening_table): E, nu  */
  DXM_HOSFORD_LINEAR = 5 /* Hosford criterion (even integer exponent a) + isotropic hardening sig0 + H p
                          * [+ (sigu-sig0)(1-exp(-b p))]: E, nu, sig0 (R0), H, a [, sigu, b] -- with linear hardening
                          * demos/multimaterials/IsotropicPlasticHosfordFlowLinear.mfront (a = 10), the matrix phase of
                          * demos/multimaterials/multimaterials.py:245-254                                          */
};

/* where a caller-supplied array lives */
enum {
  DXM_MEM_HOST = 0,    /* host (n, dim) AoS; pinned memory is DMA'd directly, pageable is staged   */
  DXM_MEM_DEVICE = 1,  /* device (n, dim) AoS on the handle's device                               */
  DXM_MEM_RESIDENT = 2 /* the handle's own SoA buffers (zero copy); pointers are ignored           */
};

typedef struct dxm_stats {
  int64_t n_points;     /* points processed by the call                                           */
  int64_t n_plastic;    /* active set size (f_trial > 0)                                          */
  int64_t n_fail;       /* local solves that hit the iteration cap or produced non-finite values   */
  int64_t max_iter;     /* maximum local Newton iteration count                                   */
  double max_residual;  /* max |r| of the local solve at exit                                     */
  double kernel_ms;     /* device time of the constitutive kernel(s) of this call (CUDA events)   */
} dxm_stats;

/* life cycle -- replaces Material.set_data_manager(ngauss) (generic.py:172-174, jaxmat.py:195-197) */
int dxm_create(int behaviour, int device, int64_t n, dxm_handle** out);
int dxm_destroy(dxm_handle* h);
int dxm_device_count(void); /* CUDA devices visible to this process (0 when there is none or no driver) */
int dxm_set_stream(dxm_handle* h, void* cuda_stream); /* run on the caller's stream (default: own) */
int64_t dxm_ld(const dxm_handle* h);                  /* SoA leading dimension (>= n)              */
int64_t dxm_npoints(const dxm_handle* h);

/* material properties -- replaces Material.update_material_property(name, values)
 * (generic.py:119-120, called from quadrature_map.py:160-172 with a 0-d or per-point array).
 * count is 1 (uniform) or n (per Gauss point). names: "E","nu","sig0","H","sigu","b"; DXM_HOSFORD_LINEAR also takes
 * "a" (uniform only: an even integer in [2, 64], default 10). */
int dxm_set_property(dxm_handle* h, const char* name, const double* v, int64_t count, int mem);

/* Piecewise-linear isotropic hardening sigma_Y(p) through the points (p[k], sig[k]), k < count (2 <= count <= 64,
 * p[0] = 0, p strictly increasing), continued with the last slope -- the device-side stand-in for the arbitrary
 * `yield_stress` callable jaxmat's vonMisesIsotropicHardening accepts (a Python callable cannot cross a C ABI; the
 * host wrapper samples it).  DXM_J2_TABLE only.  The return map walks the segments and is exact (no local Newton);
 * the per-point iteration count reports the number of segment crossings. */
int dxm_set_hardening_table(dxm_handle* h, const double* p, const double* sig, int count);

/* state -- replaces set_initial_state_dict / get_initial_state_dict / get_final_state_dict
 * (generic.py:194-201; MaterialStateManager.set_item / __getitem__, generic.py:260-292).
 * gen 0 = s0 (start of step), gen 1 = s1 (last integrate).  field: small strain
 * "strain","stress","p","epsp"; finite strain "F","PK1","p","be_bar"; also "Ct" (gen ignored). */
int dxm_field_dim(const dxm_handle* h, const char* field); /* components, <0 if unknown          */
int dxm_set_state(dxm_handle* h, int gen, const char* field, const double* v, int mem);
int dxm_get_state(dxm_handle* h, int gen, const char* field, double* out, int mem);
/* raw SoA device pointer of a field (gen 1 "strain"/"F" is where a resident caller writes the
 * gradients before dxm_integrate(..., DXM_MEM_RESIDENT, ...)); invalidated by dxm_update */
int dxm_device_ptr(dxm_handle* h, int gen, const char* field, double** ptr);
/* DLPack view (shape (dim, n), strides (ld, 1), kDLCUDA float64) of the same buffer; *out is a
 * DLManagedTensor* whose deleter drops a reference on the handle */
int dxm_export_dlpack(dxm_handle* h, int gen, const char* field, void** out);

/* the hot call -- replaces Material.integrate(gradients, dt) -> (flux, isv, Ct)
 * (generic.py:176-189, jaxmat.py:208-234), reading s0 and writing s1.
 *  grad : (n, ngrad) gradients in `mem`;  flux (n, nflux), isv (n, nisv), ct (n, nflux*ngrad) in
 *  `out_mem`; any output pointer may be NULL (not transferred).  With DXM_MEM_RESIDENT the
 *  gradients are read from s1's gradient buffer and results stay in the SoA buffers. */
int dxm_integrate(dxm_handle* h, const double* grad, int mem, double dt, double* flux, double* isv,
                  double* ct, int out_mem, dxm_stats* stats);

/* The same update restricted to the points [start, start + count) of the handle (start even); grad / flux / isv / ct hold
 * `count` rows.  For callers that overlap their own host work with the device (QuadratureExchange pipelines the
 * gather / scatter of a cell-subset map against the transfers this way).  Takes host arrays, or resident gradients
 * without outputs.  s1 is complete only once every range of the step has been integrated: do that before dxm_update or
 * reading s1.  Statistics are those of the range. */
int dxm_integrate_range(dxm_handle* h, int64_t start, int64_t count, const double* grad, int mem, double dt,
                        double* flux, double* isv, double* ct, int out_mem, dxm_stats* stats);

/* statistics of the last dxm_integrate; use it after a call made with stats == NULL, which returns without waiting for
 * the device.  The update kernel folds and publishes the record itself into page-locked mapped memory: reading it is a
 * spin on one word, with no stream synchronisation, copy or memset on the host side of a call (this fused check replaces
 * the three host NaN scans of quadrature_map.py:322-324). */
int dxm_last_stats(dxm_handle* h, dxm_stats* stats);

/* kernel_ms of dxm_stats costs two event records per launch: mode -1 = only for batches >= 262144 points (default),
 * 0 = never, 1 = always */
int dxm_enable_timing(dxm_handle* h, int mode);

/* Multi-GPU statistics: one process per GPU, Gauss points sharded by contiguous cell blocks, no exchange in
 * the update itself.  The library owns one NCCL communicator per process, used only to all-gather the 64-byte statistics
 * record of each rank ON THE HANDLE'S STREAM right after the update kernel (SUM of failed / plastic counts, MAX of
 * iterations / residual, folded by a one-warp kernel into the mapped host record): no host synchronisation is added.
 * Replaces what `MPI.COMM_WORLD.allreduce` of a failure flag would do around QuadratureMap.update under dolfinx.
 *   dxm_comm_unique_id : rank 0 fills a 128-byte id, the caller broadcasts it (torch.distributed / MPI_Bcast)
 *   dxm_comm_init      : collective over all ranks; device = this rank's GPU
 *   dxm_use_global_stats(h, 1) : this handle's dxm_stats become global (every rank must then call dxm_integrate on it)
 * On one node the records do not even go through NCCL: dxm_comm_p2p_handle (64-byte cudaIpc handle of this rank's
 * exchange buffer) -> the caller gathers the handles of all ranks, rank order -> dxm_comm_p2p_connect maps them; the
 * update kernel's publishing CTA then stores its record into every peer's buffer over NVLink, waits for the peers'
 * records and folds them itself -- no collective call, no extra launch.  If any rank cannot map its peers, call
 * dxm_comm_p2p_disable on all ranks: NCCL stays the fallback. */
int dxm_comm_p2p_handle(void* handle64);
int dxm_comm_p2p_connect(const void* handles /* nranks x 64 bytes */);
int dxm_comm_p2p_enabled(void);
int dxm_comm_p2p_disable(void);
int dxm_comm_unique_id(void* id128);
int dxm_comm_init(const void* id128, int rank, int nranks, int device);
int dxm_comm_size(void);
int dxm_comm_rank(void);
int dxm_comm_destroy(void);
int dxm_use_global_stats(dxm_handle* h, int on);

/* DataManager.update() / revert() (generic.py:212-216, jaxmat.py:39-43): O(1) generation swap */
int dxm_update(dxm_handle* h);
int dxm_revert(dxm_handle* h);

/* per-point diagnostics for parity tests: active-set flag, local iteration count, final residual */
int dxm_enable_diagnostics(dxm_handle* h, int on);
int dxm_get_diagnostics(dxm_handle* h, uint8_t* flag, int32_t* n_iter, double* resid, uint8_t* fail);

/* synthetic gradient histories written straight into s1's gradient buffer (bench / tests);
 * bit-identical to oracle/synth.py.  recipe 0 = strain (6), 1 = deformation gradient (9);
 * `start` = global index of this handle's first point (multi-GPU shards) */
int dxm_synth_gradients(dxm_handle* h, int recipe, uint64_t seed, double amp, int k, int K,
                        int64_t start);

/* GPU gradient evaluation for affine simplex meshes -- replaces
 * QuadratureExpression.eval (quadrature_function.py:45-51) + get_gradient_vals (quadrature_map.py:251-253) for the
 * registered expressions of the hot-path demos: kind 0 = Mandel vector of sym(grad u) (utils.py:146-165),
 * kind 1 = 9-vector of I + grad u (utils.py:168-190).  The mesh object holds device copies of
 *   coords (num_nodes,3) [mesh.geometry.x], geom_dofmap (num_cells,tdim+1) [mesh.geometry.dofmap],
 *   u_dofmap (num_cells,ndofs_cell) [V.dofmap.list], dphi (nqp,ndofs_cell,tdim) [basix tabulate, derivative 1].
 * dxm_eval_gradient copies the blocked displacement vector u (num_dofs*tdim) and writes the gradients of all
 * num_cells*nqp points straight into the material's gradient buffer; follow with
 * dxm_integrate(h, NULL, DXM_MEM_RESIDENT, ...). */
typedef struct dxm_mesh dxm_mesh;
int dxm_mesh_create(int device, int tdim, int64_t num_cells, int64_t num_nodes, const double* coords,
                    const int32_t* geom_dofmap, int ndofs_cell, const int32_t* u_dofmap, int64_t num_dofs,
                    int nqp, const double* dphi, dxm_mesh** out);
int dxm_mesh_destroy(dxm_mesh* m);
int dxm_eval_gradient(dxm_mesh* m, dxm_handle* h, const double* u, int mem, int kind);

/* Fused flux / tangent -> element residual / stiffness contraction -- replaces, for the same
 * affine-simplex / blocked-Lagrange setting as dxm_eval_gradient, the FFCx cell kernels DOLFINx runs over the
 * Quadrature Functions that QuadratureMap.update fills (quadrature_map.py:331-334):
 *   residual  Res = dot(flux, dgrad(v)) * qmap.dx          (assemble_vector, solvers.py:80-81)
 *   tangent   Jac = qmap.derivative(Res, u, du)            (quadrature_map.py:132-158, layout :94-104)
 * reading flux and Ct of the last dxm_integrate where they lie in HBM.  weights: reference-cell quadrature weights
 * (basix.make_quadrature, quadrature_map.py:239-243).
 * dxm_element_forms writes the element vectors fe (num_cells, nd*tdim) and matrices ke (num_cells, nd*tdim, nd*tdim)
 * (row/col = a*tdim + r, the blocked local dof order; what MatSetValuesLocal / VecSetValuesLocal take); either
 * pointer may be NULL; mem = DXM_MEM_HOST | DXM_MEM_DEVICE. */
int dxm_mesh_set_weights(dxm_mesh* m, const double* weights);
int dxm_element_forms(dxm_mesh* m, dxm_handle* h, int kind, double* fe, double* ke, int mem);

/* Device-resident assembled system: CSR pattern of the blocked space (rowptr int64 [nrows+1], colidx int32 sorted
 * within each row -- DOLFINx create_matrix / PETSc MatGetRowIJ), value array, right-hand side, optional Dirichlet
 * marker (uint8 per global dof: constrained rows and columns receive no contribution, the diagonal is set to 1,
 * the rhs entry to 0 -- the assemble_matrix(A, a, bcs) / set_bc convention for a Newton correction).
 * dxm_assemble zeroes and fills values and/or rhs from the last dxm_integrate with fp64 atomics: the tangent
 * (36 | 81 doubles per point) never leaves the device, only the assembled system does (dxm_system_get). */
typedef struct dxm_system dxm_system;
int dxm_system_create(int device, int64_t nrows, const int64_t* rowptr, const int32_t* colidx, dxm_system** out);
int dxm_system_destroy(dxm_system* s);
int dxm_system_set_bc(dxm_system* s, const uint8_t* marker); /* host array of nrows entries, NULL clears */
/* prescribed solution values x_bc on the constrained dofs (host array of nrows entries, read where marker != 0;
 * NULL: homogeneous).  dxm_assemble then moves the constrained columns to the right-hand side,
 * rhs -= A[:, bc] x_bc, and sets rhs[bc] = x_bc (DOLFINx apply_lifting + set_bc, solvers.py:84-96) */
int dxm_system_set_lifting(dxm_system* s, const double* values);
/* want_matrix == 0: residual only (the SNES function evaluation, solvers.py:80-81) -- its own kernel, the matrix values of
 * the previous pass are left alone; not available while lifting values are set (the lifting needs the matrix pass). */
int dxm_assemble(dxm_mesh* m, dxm_handle* h, int kind, dxm_system* s, int want_vector, int want_matrix);
int dxm_system_get(dxm_system* s, double* values, double* rhs, int mem); /* either may be NULL */
/* Rank-sharded assembly (one process per GPU, each holding a contiguous block of cells): every rank
 * assembles its cells into a system of the full pattern with the constrained rows deferred, the value / rhs arrays
 * (device pointers below) are summed across ranks (NCCL all-reduce -- what PETSc's MatAssembly / ghost update do for
 * the reference, solvers.py:84-96), then dxm_system_apply_constraints sets the unit diagonal and rhs[bc] once. */
int dxm_system_defer_constraints(dxm_system* s, int on);
int dxm_system_apply_constraints(dxm_system* s);
int dxm_system_device_ptrs(dxm_system* s, double** values, double** rhs);
int64_t dxm_system_nnz(const dxm_system* s);
/* Device-resident Krylov solve A x = rhs of the assembled system (BiCGStab, block x block Jacobi preconditioner,
 * x0 = 0, stop at |r| <= rtol |rhs|); returns 0 converged, 1 maxit reached, < 0 error.  NOT part of the drop-in
 * path -- the reference hands the system to PETSc (solvers.py:182-196); it exists so that the config-5 Newton loop
 * can run in a DOLFINx/PETSc-free environment (scripts/newton_bar.py). */
int dxm_system_solve(dxm_system* s, int block, double rtol, int maxit, double* x, int mem, int* iters,
                     double* relres);

/* pinned host memory helpers (the Python wrapper allocates its output arrays with these) */
int dxm_host_alloc(void** ptr, int64_t bytes);
int dxm_host_free(void* ptr);
/* page-lock / release a caller-owned host array in place (e.g. a dolfinx Function's x.array), so that
 * dxm_integrate / dxm_get_state DMA straight into it -- the "contiguous-range fast path" that replaces the
 * fancy-index gather/scatter of utils.py:98-143 when a QuadratureMap covers all cells */
int dxm_host_register(void* ptr, int64_t bytes);
int dxm_host_unregister(void* ptr);

/* Host half of the packed-tangent hand-off: expands packed (n, 21) rows of the symmetric 6x6 tangent (row map in the
 * conventions above) into the reference's row-major (n, 36) array (quadrature_map.py:334) on `threads` host threads
 * (<= 0: library default, DXM_HOST_THREADS).  Pure data movement, bit for bit.  dxm_integrate(..., DXM_MEM_HOST, ...)
 * uses it internally when DXM_HOST_MIRROR=1: the device then sends 21 instead of 36 doubles per point over PCIe and
 * the mirror runs on the host while the next chunk is in flight. */
int dxm_host_mirror_sym6(const double* packed, double* full, int64_t n, int threads);

/* Host-side row gather / scatter by index on the library's thread pool (pure data movement, no device involved):
 *   gather : dst[r, :] = src[rows[r], :]        scatter : dst[rows[r], :] = src[r, :]        rows of row_len doubles
 * -- the two passes a QuadratureMap built on a cell SUBSET needs around integrate(), because the Function arrays
 * span the whole mesh: _get_vals(gradient)[dofs, :] (quadrature_map.py:251-253) and fun.x.array[dofs] = arr
 * (utils.py:136-143).  rows must be distinct for scatter (they are: one entry per cell). threads <= 0: pool default. */
int dxm_host_gather_rows(const double* src, const int64_t* rows, int64_t n, int64_t row_len, double* dst, int threads);
int dxm_host_scatter_rows(double* dst, const int64_t* rows, int64_t n, int64_t row_len, const double* src, int threads);

/* measurement support */
int64_t dxm_launch_count(void);                       /* kernels launched by this library so far */
int dxm_fp64_peak(int device, double* tflops);        /* register-resident DFMA microbenchmark   */
int dxm_copy_peak(int device, int64_t bytes, double* gbs); /* device copy bandwidth (read+write)  */
/* pure-traffic twin of the update kernels: nread coalesced read streams + nwrite write streams over n
 * points (supported mixes 25/49 = J2 with a full tangent, 25/34 = J2 with the packed tangent, 25/97 = FeFp, 37/37, 1/1): the practical HBM ceiling for that mix */
int dxm_stream_peak(int device, int64_t n, int nread, int nwrite, double* gbs);

const char* dxm_last_error(void);
const char* dxm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DXM_H */
