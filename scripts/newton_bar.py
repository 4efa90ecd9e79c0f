"""Config 5 surrogate: a displacement-driven Newton solve of a 3-D P2 notched bar with the finite-strain FeFp
behaviour, entirely resident on one B200 -- u -> gradients (dxm_eval_gradient) -> constitutive update
(dxm_integrate) -> fused element contraction + CSR assembly (dxm_assemble) -> Krylov solve (dxm_system_solve).

It mirrors demos/jax/finite_strain_elastoplasticity/finite_strain_elastoplasticity.py of the reference (P2 tets,
quadrature degree 2 -> 4 points per cell, symmetry planes x=0 / y=0 / z=0, imposed u_x on the far face,
F = I + grad u, Res = PK1 . dF(v) dx, Jac = qmap.derivative(Res), newtonls without line search, rtol = atol = 1e-8),
with a structured Kuhn mesh instead of gmsh and BiCGStab + block-Jacobi instead of PETSc GMRES + GAMG -- DOLFINx,
PETSc, gmsh and MPI are not available in this image, so this is a surrogate of the reference's driver, not a run of it.

    python scripts/newton_bar.py [nx ny nz] [--steps S] [--strain E] [--json out.json]
    torchrun --nproc-per-node N scripts/newton_bar.py ...        # N ranks, one GPU each

With N ranks the cells are split into N contiguous blocks: each rank evaluates gradients, runs the
constitutive update and assembles for its own cells only -- no exchange, as in the reference where every MPI rank owns
a mesh partition -- then the assembled values / right-hand side are summed over the ranks with one NCCL all-reduce
(the FE side's only exchange step, what PETSc's assembly does for the reference).  The Krylov stand-in is not
distributed: rank 0 solves and broadcasts the correction.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PAIRS = ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))
QP_DEG2 = np.array([[0.1381966011250105] * 3, [0.5854101966249685, 0.1381966011250105, 0.1381966011250105],
                    [0.1381966011250105, 0.5854101966249685, 0.1381966011250105],
                    [0.1381966011250105, 0.1381966011250105, 0.5854101966249685]])
W_DEG2 = np.full(4, 1.0 / 24.0)


def p2_tet_dphi(points):
    """reference gradients (nqp, 10, 3) of the P2 Lagrange basis: 4 vertices then edge midpoints in PAIRS order"""
    pts = np.asarray(points, dtype=np.float64)
    gl = np.array([[-1.0, -1.0, -1.0], [1, 0, 0], [0, 1, 0], [0, 0, 1]])
    lam = np.stack([1 - pts.sum(1), pts[:, 0], pts[:, 1], pts[:, 2]], axis=1)
    out = np.zeros((len(pts), 10, 3))
    for a in range(4):
        out[:, a, :] = (4 * lam[:, a] - 1)[:, None] * gl[a]
    for e, (a, b) in enumerate(PAIRS):
        out[:, 4 + e, :] = 4 * (lam[:, a][:, None] * gl[b] + lam[:, b][:, None] * gl[a])
    return out


def bar_mesh(nx, ny, nz, L=10.0, W=1.0, notch=0.2):
    """Structured Kuhn mesh (6 tets per hex) of a quarter bar [0,L]x[0,W]x[0,W] whose cross-section narrows by
    `notch` around x = 0 (the symmetry plane through the notch).  P2 nodes are exactly the points of the
    (2nx+1)(2ny+1)(2nz+1) half-step grid, so the dofmap is pure index arithmetic.
    Returns vertex coords (nv,3), geometry dofmap (nc,4) int32, P2 dofmap (nc,10) int32, P2 node coords (nn,3)."""
    fx, fy, fz = 2 * nx + 1, 2 * ny + 1, 2 * nz + 1
    gx, gy, gz = np.meshgrid(np.arange(fx), np.arange(fy), np.arange(fz), indexing="ij")
    x = gx.ravel() * (L / (2 * nx))
    shrink = 1.0 - notch * np.exp(-((x / (0.08 * L)) ** 2))
    nodes = np.stack([x, gy.ravel() * (W / (2 * ny)) * shrink, gz.ravel() * (W / (2 * nz)) * shrink], axis=1)
    fid = lambda i, j, k: (i * fy + j) * fz + k  # noqa: E731  fine-grid node id
    ci, cj, ck = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    ci, cj, ck = 2 * ci.ravel(), 2 * cj.ravel(), 2 * ck.ravel()
    corner = lambda a, b, c: (ci + 2 * a, cj + 2 * b, ck + 2 * c)  # noqa: E731
    v = [corner(a, b, c) for a in (0, 1) for b in (0, 1) for c in (0, 1)]  # index 4a+2b+c
    tets = []
    for p in ((1, 3), (1, 5), (2, 3), (2, 6), (4, 5), (4, 6)):
        tets.append([v[0], v[p[0]], v[p[1]], v[7]])
    ud_cols = []
    for t in tets:
        cols = [fid(*vv) for vv in t]
        for a, b in PAIRS:
            cols.append(fid((t[a][0] + t[b][0]) // 2, (t[a][1] + t[b][1]) // 2, (t[a][2] + t[b][2]) // 2))
        ud_cols.append(np.stack(cols, axis=1))
    # cell order: hex-major, the 6 tets of a hex consecutive (contiguous Gauss-point ranges per hex block)
    ud = np.stack(ud_cols, axis=1).reshape(-1, 10).astype(np.int32)
    gd = ud[:, :4].copy()
    # affine cells: the P2 edge nodes sit at the midpoints of the (mapped) vertices
    for e, (a, b) in enumerate(PAIRS):
        nodes[ud[:, 4 + e]] = 0.5 * (nodes[ud[:, a]] + nodes[ud[:, b]])
    return nodes, gd, ud, nodes


def sparsity(ud, num_nodes, tdim=3):
    """CSR pattern of the blocked P2 space (what DOLFINx create_matrix derives from the dofmap)."""
    import scipy.sparse as sp

    nd = ud.shape[1]
    rows = np.repeat(ud, nd, axis=1).ravel()
    cols = np.tile(ud, (1, nd)).ravel()
    P = sp.csr_matrix((np.ones(len(rows), dtype=np.int8), (rows, cols)), shape=(num_nodes, num_nodes))
    P.sum_duplicates()
    P.sort_indices()
    # expand the node pattern to tdim x tdim blocks
    cnt = np.diff(P.indptr)
    rowptr = np.zeros(num_nodes * tdim + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(np.repeat(cnt * tdim, tdim))
    flat = (P.indices[:, None].astype(np.int32) * tdim + np.arange(tdim, dtype=np.int32)[None, :]).ravel()
    # blocked row (i, r) repeats node row i's column list: gather with one index vector
    row_len = np.repeat(cnt * tdim, tdim)
    shift = np.repeat(P.indptr[:-1].astype(np.int64) * tdim, tdim) - rowptr[:-1]
    src = np.repeat(shift, row_len)
    src += np.arange(rowptr[-1], dtype=np.int64)
    colidx = flat[src]
    return rowptr, colidx


def boundary_conditions(nodes, L):
    """symmetry planes x=0 (u_x), y=0 (u_y), z=0 (u_z) clamped; u_x imposed on x=L (demo :134-139)"""
    n = len(nodes)
    bc = np.zeros((n, 3), dtype=bool)
    tol = 1e-9
    bc[nodes[:, 0] < tol, 0] = True
    bc[nodes[:, 1] < tol, 1] = True
    bc[nodes[:, 2] < tol, 2] = True
    top = nodes[:, 0] > L - tol
    bc[top, 0] = True
    return bc.ravel(), np.flatnonzero(top) * 3


def run_gpu(nx, ny, nz, steps=3, strain=0.01, L=10.0, W=1.0, newton_rtol=1e-8, newton_atol=1e-8, ksp_rtol=1e-8,
            ksp_maxit=50000, max_newton=20, verbose=True, props=None):
    import torch
    import torch.distributed as dist

    import dolfinx_materials_b200 as jm
    from dolfinx_materials_b200.distributed import init_stats_comm, shard_range
    from dolfinx_materials_b200.fe import AssembledSystem, ElementForms, GradientEvaluator

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    device = torch.cuda.current_device()
    verbose = verbose and rank == 0

    props = props or dict(E=70e3, nu=0.3, sig0=500.0, sigu=750.0, b=1000.0)
    t0 = time.perf_counter()
    coords, gd, ud, nodes = bar_mesh(nx, ny, nz, L, W)
    dphi = p2_tet_dphi(QP_DEG2)
    nc_global, nqp = len(gd), 4
    c0, c1 = shard_range(nc_global, rank, world)  # this rank's contiguous block of cells
    nc = c1 - c0
    mat = jm.CUDAMaterial(jm.FeFpJ2Plasticity(
        elasticity=jm.LinearElasticIsotropic(E=props["E"], nu=props["nu"]),
        yield_stress=jm.VoceHardening(sig0=props["sig0"], sigu=props["sigu"], b=props["b"])), device=device)
    mat.set_data_manager(nc * nqp)
    if world > 1:
        # failure / active-set counts and residual maxima reduced over the ranks in-stream: the library's NCCL
        # communicator all-gathers the 64-byte record right after the update kernel (no host-side collective)
        if init_stats_comm() == world:
            mat.use_global_stats()
    mat.enable_timing(1)
    ge = GradientEvaluator(mat, coords, gd[c0:c1], ud[c0:c1], dphi, tdim=3, num_dofs=len(nodes))
    forms = ElementForms(ge, W_DEG2)
    rowptr, colidx = sparsity(ud, len(nodes))
    bc, top_dofs = boundary_conditions(nodes, L)
    system = AssembledSystem(forms, rowptr, colidx, bc=bc)
    t_setup = time.perf_counter() - t0
    u = np.zeros(3 * len(nodes))
    timers = dict(gradients=0.0, update=0.0, assemble=0.0, solve=0.0, rhs_d2h=0.0)
    history = []
    kernel_ms = 0.0
    t_loop = time.perf_counter()
    for k in range(1, steps + 1):
        # the imposed displacement increment enters through the first linear solve (apply_lifting), i.e. the
        # elastic-predictor start DOLFINx's Newton solver makes from the previous converged state
        lift = np.zeros_like(u)
        lift[top_dofs] = -(strain * L / steps)  # we solve A x = R and set u -= x
        r0 = None
        for it in range(max_newton + 1):
            t = time.perf_counter(); ge.eval(u); timers["gradients"] += time.perf_counter() - t
            t = time.perf_counter(); st = mat.integrate_resident(); timers["update"] += time.perf_counter() - t
            kernel_ms += st.kernel_ms
            if st.n_fail:
                raise RuntimeError(f"{st.n_fail} local solves failed")
            t = time.perf_counter()
            system.set_lifting(lift if it == 0 else None)
            system.assemble_sharded()  # local cells, then one NCCL all-reduce of values / rhs when world > 1
            timers["assemble"] += time.perf_counter() - t
            t = time.perf_counter(); _, rhs = system.get(values=False); timers["rhs_d2h"] += time.perf_counter() - t
            rn = float(np.linalg.norm(rhs))
            r0 = rn if r0 is None else r0
            if verbose:
                print(f"step {k} newton {it}: |R| = {rn:.6e}  plastic {st.n_plastic / (nc_global * nqp):.3f}", flush=True)
            if it > 0 and (rn <= newton_atol or rn <= newton_rtol * r0):
                break
            if it == max_newton:
                raise RuntimeError("Newton did not converge")
            t = time.perf_counter()
            if rank == 0:
                du, kit, rel, ok = system.solve(rtol=ksp_rtol, maxit=ksp_maxit)
            if world > 1:  # the stand-in solver is not distributed: rank 0 solves, everyone gets the same correction
                meta = torch.tensor([kit, rel, float(ok)] if rank == 0 else [0.0, 0.0, 0.0], dtype=torch.float64, device="cuda")
                dut = torch.from_numpy(du).cuda() if rank == 0 else torch.empty(u.size, dtype=torch.float64, device="cuda")
                dist.broadcast(meta, 0)
                dist.broadcast(dut, 0)
                du = dut.cpu().numpy()
                kit, rel, ok = int(meta[0].item()), meta[1].item(), bool(meta[2].item())
            timers["solve"] += time.perf_counter() - t
            if not ok:
                raise RuntimeError(f"Krylov solve stalled at {rel:.3e} after {kit} iterations")
            history.append(dict(step=k, newton=it, residual=rn, krylov_iterations=kit, krylov_relres=rel))
            u -= du
        mat.data_manager.update()
    t_total = time.perf_counter() - t_loop
    n_updates = len(history) + steps
    # per-update wall time of the resident call (kernel + in-stream statistics all-gather + publish) without the timing
    # events, next to the kernel's own time with them
    mat.data_manager.revert()
    mat.enable_timing(0)
    for _ in range(20):
        mat.integrate_resident()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(200):
        mat.integrate_resident()
    update_wall_us = (time.perf_counter() - t) / 200 * 1e6
    mat.enable_timing(1)
    update_kernel_us = min(mat.integrate_resident().kernel_ms for _ in range(20)) * 1e3
    if world > 1:
        tt = torch.tensor([update_wall_us, update_kernel_us], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        update_wall_us, update_kernel_us = tt.tolist()
    if world > 1:  # per-stage times: slowest rank
        tt = torch.tensor([timers[k] for k in sorted(timers)] + [kernel_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        for k, v in zip(sorted(timers), tt.tolist()):
            timers[k] = v
        kernel_ms = tt[-1].item()
    nc = nc_global
    info = dict(ranks=world, cells=nc, points=nc * nqp, dofs=u.size, nnz=int(system.nnz), steps=steps, strain=strain,
                newton_iterations=len(history), constitutive_updates=n_updates, setup_s=t_setup, loop_s=t_total,
                timers_s=timers, update_kernel_ms_total=kernel_ms, update_wall_us=update_wall_us,
                update_kernel_us=update_kernel_us,
                constitutive_share=timers["update"] / t_total,
                update_gps=nc * nqp * n_updates / max(timers["update"], 1e-12),
                krylov_iterations_total=int(sum(h["krylov_iterations"] for h in history)),
                plastic_fraction=st.n_plastic / (nc * nqp),
                u_l2=float(np.linalg.norm(u)), u_probe=[float(x) for x in u[:: max(1, u.size // 7)][:7]])
    return u, mat, info, history


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("n", nargs="*", type=int, default=[40, 8, 8])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--strain", type=float, default=0.01)
    ap.add_argument("--json", default=None)
    ap.add_argument("--ksp-rtol", type=float, default=1e-8)
    a = ap.parse_args()
    nx, ny, nz = (a.n + [8, 8])[:3]
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    u, mat, info, hist = run_gpu(nx, ny, nz, steps=a.steps, strain=a.strain, ksp_rtol=a.ksp_rtol)
    if world > 1:
        rank0 = dist.get_rank() == 0
        dist.barrier()
        del mat
        from dolfinx_materials_b200 import _lib
        _lib.load().dxm_comm_destroy()
        dist.destroy_process_group()
        if not rank0:
            sys.exit(0)
    print(json.dumps(info, indent=1))
    if a.json:
        os.makedirs(os.path.dirname(a.json) or ".", exist_ok=True)
        json.dump(dict(info=info, history=hist), open(a.json, "w"), indent=1)
