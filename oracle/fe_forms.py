"""numpy restatement of the FE-side contraction that consumes the constitutive update's outputs -- SURVEY.md
section 8(f) rank 3.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Reference path: after ``QuadratureMap.update`` has written flux and flattened tangent into their Quadrature
Functions (``dolfinx_materials/quadrature_map.py:331-334``), DOLFINx assembles

* the residual ``Res = dot(flux, dgrad(v)) * qmap.dx`` (``solvers.py:80-81``; forms e.g.
  ``demos/jax/finite_strain_elastoplasticity/finite_strain_elastoplasticity.py:171-172``,
  ``tests/uniaxial_tension.py:57-60``) and
* the tangent ``Jac = qmap.derivative(Res, u, du)`` = ``dgrad(v) . Ct . dgrad(du) * dx``
  (``quadrature_map.py:132-158`` with ``jacobians[...]`` the row-major view of ``jacobian_flatten``,
  ``quadrature_map.py:94-104``)

cell by cell with FFCx-generated kernels.  For the hot-path gradients (kind 0: Mandel vector of ``sym(grad u)``,
kind 1: 9-vector of ``I + grad u``) on affine simplices with a blocked Lagrange space these kernels are, with
``g[a, j]`` the physical gradient of basis function ``a`` and ``vol_q = w_q |det J|``:

    fe[(a, r)]         = sum_q vol_q  sum_j   S_rj(q) g[a, j]
    ke[(a, r), (b, s)] = sum_q vol_q  sum_jl  g[a, j] A_(rj)(sl)(q) g[b, l]

where ``S`` is the flux as a tensor and ``A = dS/dG`` the tangent as a 4th-order tensor.  For kind 1 these are
the 9-vector entries themselves; for kind 0 the Mandel factors are undone: ``S_rj = sigma_M[m(rj)] * f(rj)``,
``A = Ct[m(rj), m(sl)] * f(rj) f(sl)`` with ``f = 1`` on the diagonal and ``1/sqrt(2)`` off it
(``utils.py:146-165``).  Local dof ``(a, r)`` has index ``a * tdim + r`` (blocked space), point ordering is the
reference's ``num_qp * cell + q`` (``quadrature_map.py:255-260``).

PARITY STATUS: **unpinned against DOLFINx** -- dolfinx / ffcx / basix are not installable in the build container, so no
reference-executed golden vectors exist for this row.  The restatement is checked instead against an independent
dense B-matrix / einsum formulation, the patch test, linearity for an elastic material (fe == ke u_e) and a
finite-difference tangent through the FeFp update (``tests/test_oracle_fe_forms.py``), and end to end by the Newton
loop of ``tests/test_newton_bar_gpu.py`` (quadratic convergence only happens with a consistent residual / tangent pair).

Operation order is canonical (explicit loops, no einsum, fused multiply-adds placed by hand: ``oracle/canon.py``) and
shared with ``fe_forms_kernel``; element vectors and matrices therefore agree bit for bit.  Global assembly sums element contributions (order-dependent on the GPU:
atomics), compared with a tolerance.
"""

import numpy as np

RSQRT2 = 0.70710678118654752440


def idx9(i, j):
    """position of (i, j) in [11,22,33,12,21,13,31,23,32] (utils.py:173-186)"""
    if i == j:
        return i
    return {(0, 1): 3, (1, 0): 4, (0, 2): 5, (2, 0): 6, (1, 2): 7, (2, 1): 8}[(i, j)]


def idx6(i, j):
    """Mandel position of (i, j) in [11,22,33,12,13,23] (utils.py:151-161)"""
    if i == j:
        return i
    return {(0, 1): 3, (0, 2): 4, (1, 2): 5}[(min(i, j), max(i, j))]


def geometry(coords, geom_dofmap, tdim):
    """K = J^-1 (list of lists of per-cell arrays) and det J, same order as oracle.fe_gradient.evaluate"""
    coords = np.asarray(coords, dtype=np.float64)
    gd = np.asarray(geom_dofmap)
    x = [coords[gd[:, v]] for v in range(tdim + 1)]
    J = [[x[j + 1][:, i] - x[0][:, i] for j in range(tdim)] for i in range(tdim)]
    if tdim == 2:
        det = J[0][0] * J[1][1] - J[0][1] * J[1][0]
        rdet = 1.0 / det
        K = [[J[1][1] * rdet, -(J[0][1] * rdet)], [-(J[1][0] * rdet), J[0][0] * rdet]]
    else:
        c = [[None] * 3 for _ in range(3)]
        c[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1]
        c[0][1] = J[0][2] * J[2][1] - J[0][1] * J[2][2]
        c[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1]
        c[1][0] = J[1][2] * J[2][0] - J[1][0] * J[2][2]
        c[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0]
        c[1][2] = J[0][2] * J[1][0] - J[0][0] * J[1][2]
        c[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0]
        c[2][1] = J[0][1] * J[2][0] - J[0][0] * J[2][1]
        c[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0]
        det = (J[0][0] * c[0][0] + J[0][1] * c[1][0]) + J[0][2] * c[2][0]
        rdet = 1.0 / det
        K = [[c[i][j] * rdet for j in range(3)] for i in range(3)]
    return K, det


def _flux_tensor(flux, kind, q_pts, r, j):
    if kind == 1:
        return flux[q_pts, idx9(r, j)]
    v = flux[q_pts, idx6(r, j)]
    return v if r == j else v * RSQRT2


def _tangent_tensor(ct, kind, q_pts, r, j, s, l):
    if kind == 1:
        return ct[q_pts, idx9(r, j) * 9 + idx9(s, l)]
    v = ct[q_pts, idx6(r, j) * 6 + idx6(s, l)]
    noff = (r != j) + (s != l)
    return v if noff == 0 else (v * RSQRT2 if noff == 1 else v * 0.5)


def element_forms(coords, geom_dofmap, u_dofmap, dphi, weights, flux, ct, kind, tdim, want_matrix=True):
    """flux (n, 6|9), ct (n, 36|81) or (n, nf, ng) with n = ncells*nqp.  Returns fe (ncells, nd*tdim) and
    ke (ncells, nd*tdim, nd*tdim) (None if not wanted).

    Canonical operation order (shared with ``fe_forms_kernel``, one warp lane per column ``(b, s)``):
    ``gv_q[a][j] = vol_q g_q[a][j]``; ``U_q[(r,j)] = sum_l A_q[(r,j)][(s,l)] g_q[b][l]``;
    ``Ut_q[(r,m)] = vol_q sum_j K[m][j] U_q[(r,j)]`` (U carried back to the reference cell: on an affine simplex
    ``g_q[a][j] = sum_m dphi_q[a][m] K[m][j]`` with one ``K`` per cell, so the row operand of an entry is the tabulated
    reference gradient -- the same numbers for every cell, constants of the kernel);
    ``ke[(a,r),(b,s)] = sum_q sum_m dphi_q[a][m] Ut_q[(r,m)]`` and ``fe[(b,s)] = sum_q sum_j S_q[s][j] gv_q[b][j]``, every
    sum one product followed by fused multiply-adds in the order written (q outer, j / l / m inner)."""
    from .canon import fma

    ud = np.asarray(u_dofmap)
    dphi = np.asarray(dphi, dtype=np.float64)
    weights = np.asarray(weights, dtype=np.float64)
    nc, nd = ud.shape
    nqp = dphi.shape[0]
    flux = np.asarray(flux, dtype=np.float64)
    ct = np.asarray(ct, dtype=np.float64).reshape(nc * nqp, -1)
    K, det = geometry(coords, geom_dofmap, tdim)
    adet = np.abs(det)
    ndof = nd * tdim
    fe = np.zeros((nc, ndof))
    ke = np.zeros((nc, ndof, ndof)) if want_matrix else None
    cells = np.arange(nc)
    g, gv, pts = [], [], []
    for q in range(nqp):
        pts.append(cells * nqp + q)
        vol = weights[q] * adet
        gq = [[None] * tdim for _ in range(nd)]
        for a in range(nd):
            for j in range(tdim):
                acc = dphi[q, a, 0] * K[0][j]
                for m in range(1, tdim):
                    acc = acc + dphi[q, a, m] * K[m][j]
                gq[a][j] = acc
        g.append(gq)
        gv.append([[vol * gq[a][j] for j in range(tdim)] for a in range(nd)])
    for b in range(nd):
        for s in range(tdim):
            col = b * tdim + s
            acc = None
            for q in range(nqp):
                for j in range(tdim):
                    sv = _flux_tensor(flux, kind, pts[q], s, j)
                    acc = sv * gv[q][b][0] if acc is None else fma(sv, gv[q][b][j], acc)
            fe[:, col] = acc
            if not want_matrix:
                continue
            U = [[[None] * tdim for _ in range(tdim)] for _ in range(nqp)]
            for q in range(nqp):
                for r in range(tdim):
                    for j in range(tdim):
                        u = _tangent_tensor(ct, kind, pts[q], r, j, s, 0) * g[q][b][0]
                        for l in range(1, tdim):
                            u = fma(_tangent_tensor(ct, kind, pts[q], r, j, s, l), g[q][b][l], u)
                        U[q][r][j] = u
            Ut = [[[None] * tdim for _ in range(tdim)] for _ in range(nqp)]
            for q in range(nqp):
                vol = weights[q] * adet
                for r in range(tdim):
                    for m in range(tdim):
                        t = K[m][0] * U[q][r][0]
                        for j in range(1, tdim):
                            t = fma(K[m][j], U[q][r][j], t)
                        Ut[q][r][m] = vol * t
            for a in range(nd):
                for r in range(tdim):
                    acc = None
                    for q in range(nqp):
                        for m in range(tdim):
                            acc = dphi[q, a, 0] * Ut[q][r][0] if acc is None else fma(dphi[q, a, m], Ut[q][r][m], acc)
                    ke[:, a * tdim + r, col] = acc
    return fe, ke


def global_dofs(u_dofmap, tdim):
    ud = np.asarray(u_dofmap)
    return (ud[:, :, None] * tdim + np.arange(tdim)[None, None, :]).reshape(len(ud), -1)


def sparsity(u_dofmap, num_dofs, tdim):
    """CSR pattern (rowptr int64, colidx int32, sorted columns) of the blocked space, as DOLFINx's
    ``create_matrix`` builds it from the dofmap."""
    import scipy.sparse as sp

    gdofs = global_dofs(u_dofmap, tdim)
    nd = gdofs.shape[1]
    rows = np.repeat(gdofs, nd, axis=1).ravel()
    cols = np.tile(gdofs, (1, nd)).ravel()
    n = num_dofs * tdim
    P = sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(n, n))
    P.sum_duplicates()
    P.sort_indices()
    return P.indptr.astype(np.int64), P.indices.astype(np.int32)


def assemble(u_dofmap, fe, ke, num_dofs, tdim, bc=None, lift=None, constrain=True):
    """Global residual vector and CSR tangent from the element forms.  ``bc``: optional bool marker per global
    dof -- constrained rows/columns receive no contribution and a unit diagonal (what
    ``assemble_matrix(A, a, bcs)`` + ``set_bc`` leave for a Newton correction with homogeneous increments).
    ``lift``: prescribed solution values on the constrained dofs -- ``b -= A[:, bc] lift[bc]`` on the free rows
    and ``b[bc] = lift[bc]`` (``apply_lifting`` + ``set_bc``, ``solvers.py:84-96``).  ``constrain=False`` leaves the
    constrained rows empty (a rank's partial contribution, to be summed across ranks before ``apply_constraints``)."""
    import scipy.sparse as sp

    gdofs = global_dofs(u_dofmap, tdim)
    n = num_dofs * tdim
    b = np.zeros(n)
    free = np.ones(n, dtype=bool) if bc is None else ~np.asarray(bc, dtype=bool)
    fe_m = np.where(free[gdofs], fe, 0.0)
    np.add.at(b, gdofs.ravel(), fe_m.ravel())
    A = None
    if ke is not None:
        nd = gdofs.shape[1]
        rows = np.repeat(gdofs, nd, axis=1).ravel()
        cols = np.tile(gdofs, (1, nd)).ravel()
        if bc is not None and lift is not None:
            lift = np.asarray(lift, dtype=np.float64)
            moved = np.where(free[rows] & ~free[cols], ke.ravel() * lift[cols], 0.0)
            np.subtract.at(b, rows, moved)
            if constrain:
                b[~free] = lift[~free]
        vals = np.where(free[rows] & free[cols], ke.ravel(), 0.0)
        A = sp.csr_matrix((vals, (rows, cols)), shape=(n, n))
        A.sum_duplicates()
        if bc is not None and constrain:
            A = A + sp.diags(np.where(free, 0.0, 1.0), format="csr")
        A.sort_indices()
    return b, A


def apply_constraints(A, b, bc, lift=None):
    """Unit diagonal and prescribed rhs on the constrained dofs of a summed (rank-sharded) system."""
    import scipy.sparse as sp

    free = ~np.asarray(bc, dtype=bool)
    A = A + sp.diags(np.where(free, 0.0, 1.0), format="csr")
    b = b.copy()
    b[~free] = 0.0 if lift is None else np.asarray(lift)[~free]
    return A, b
