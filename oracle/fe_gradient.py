"""numpy restatement of the FE-side gradient evaluation that feeds ``Material.integrate`` -- SURVEY.md
section 8(f) rank 2.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Reference path: ``QuadratureExpression.eval`` (``dolfinx_materials/quadrature_function.py:45-51``) evaluates a
compiled ``fem.Expression`` of the registered UFL gradient at every quadrature point of every cell and
``QuadratureMap.get_gradient_vals`` gathers it (``quadrature_map.py:251-253``).  For the hot-path behaviours
the registered expressions are (demos):

* ``strain(u)`` = Mandel vector of ``sym(grad u)`` (``utils.py:146-165``; plane problems pad with zeros,
  ``demos/jax/elastoplasticity/plane_elastoplasticity.py:118-128``)                       -> kind 0
* ``F(u) = nonsymmetric_tensor_to_vector(Id + grad(u))`` (``utils.py:168-190``,
  ``demos/jax/finite_strain_elastoplasticity/finite_strain_elastoplasticity.py:143-147``)  -> kind 1

on affine simplex meshes with a blocked Lagrange displacement space.  With ``K = J^-1`` (constant per affine
cell) and the reference-element basis gradients tabulated at the quadrature points ``dphi[q, a, :]`` (what
``basix`` tabulates; passed in, so any dof numbering works):

    H[r, j] = sum_a u[bs * dofmap[c, a] + r] * dphi[q, a, j]        grad u = H K

Point ordering is the reference's: point ``num_qp * cell + q`` (``quadrature_map.py:255-260``).
Operation order is canonical (explicit loops, no einsum) and shared with ``fe_gradient_kernel``.
"""

import numpy as np

RSQRT2 = 0.70710678118654752440


def evaluate(coords, geom_dofmap, u_dofmap, u, dphi, kind, tdim):
    """coords (nnodes, 3); geom_dofmap (ncells, tdim+1); u_dofmap (ncells, nd); u flat (ndofs*tdim,) blocked;
    dphi (nqp, nd, tdim).  Returns (ncells*nqp, 6) for kind 0, (ncells*nqp, 9) for kind 1."""
    coords = np.asarray(coords, dtype=np.float64)
    gd = np.asarray(geom_dofmap)
    ud = np.asarray(u_dofmap)
    u = np.asarray(u, dtype=np.float64).reshape(-1, tdim)
    dphi = np.asarray(dphi, dtype=np.float64)
    nc, nd = ud.shape
    nqp = dphi.shape[0]
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
    ua = [[u[ud[:, a], r] for r in range(tdim)] for a in range(nd)]
    ncomp = 6 if kind == 0 else 9
    out = np.zeros((nc, nqp, ncomp))
    for q in range(nqp):
        H = [[ua[0][r] * dphi[q, 0, j] for j in range(tdim)] for r in range(tdim)]
        for a in range(1, nd):
            for r in range(tdim):
                for j in range(tdim):
                    H[r][j] = H[r][j] + ua[a][r] * dphi[q, a, j]
        G = [[0.0] * 3 for _ in range(3)]
        for r in range(tdim):
            for i in range(tdim):
                acc = H[r][0] * K[0][i]
                for j in range(1, tdim):
                    acc = acc + H[r][j] * K[j][i]
                G[r][i] = acc
        if kind == 0:
            vals = [G[0][0], G[1][1], G[2][2], (G[0][1] + G[1][0]) * RSQRT2, (G[0][2] + G[2][0]) * RSQRT2,
                    (G[1][2] + G[2][1]) * RSQRT2]
        else:
            vals = [1.0 + G[0][0], 1.0 + G[1][1], 1.0 + G[2][2], G[0][1], G[1][0], G[0][2], G[2][0], G[1][2], G[2][1]]
        for k, v in enumerate(vals):
            out[:, q, k] = v
    return out.reshape(nc * nqp, ncomp)


# ---- small self-contained mesh / element helpers for the tests (not part of the restated path) -----------
def box_tets(nx, ny, nz, order=1):
    """Structured tetrahedral mesh of the unit cube (6 tets per hex); returns coords (n,3), geometry dofmap
    (ncells,4), displacement dofmap (ncells, 4|10) and the coordinates of the displacement nodes."""
    xs, ys, zs = np.linspace(0, 1, nx + 1), np.linspace(0, 1, ny + 1), np.linspace(0, 1, nz + 1)
    X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")
    coords = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    # a mild smooth distortion so that Jacobians differ from cell to cell (cells stay affine)
    coords = coords + 0.03 * np.stack([np.sin(3 * coords[:, 1]), np.sin(2 * coords[:, 2]), np.sin(4 * coords[:, 0])], axis=1)
    nid = lambda i, j, k: (i * (ny + 1) + j) * (nz + 1) + k  # noqa: E731
    tets = []
    for i in range(nx):
        for j in range(ny):
            for k in range(nz):
                v = [nid(i + a, j + b, k + c) for a in (0, 1) for b in (0, 1) for c in (0, 1)]
                # v index = 4a + 2b + c ; Kuhn subdivision along the diagonal v0-v7
                for p in ((1, 3), (1, 5), (2, 3), (2, 6), (4, 5), (4, 6)):
                    tets.append([v[0], v[p[0]], v[p[1]], v[7]])
    gd = np.array(tets, dtype=np.int32)
    if order == 1:
        return coords, gd, gd.copy(), coords.copy()
    edges = {}
    nodes = [c for c in coords]
    ud = np.zeros((len(gd), 10), dtype=np.int32)
    ud[:, :4] = gd
    pairs = ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))
    for c, t in enumerate(gd):
        for e, (a, b) in enumerate(pairs):
            key = (min(t[a], t[b]), max(t[a], t[b]))
            if key not in edges:
                edges[key] = len(nodes)
                nodes.append(0.5 * (coords[t[a]] + coords[t[b]]))
            ud[c, 4 + e] = edges[key]
    return coords, gd, ud, np.array(nodes)


def tet_dphi(points, order):
    """Reference gradients (nqp, nd, 3) of the P1 / P2 Lagrange basis on the unit tetrahedron, node order:
    4 vertices, then edge midpoints (0,1),(0,2),(0,3),(1,2),(1,3),(2,3) (matches ``box_tets``)."""
    pts = np.asarray(points, dtype=np.float64)
    gl = np.array([[-1.0, -1.0, -1.0], [1, 0, 0], [0, 1, 0], [0, 0, 1]])  # grad of barycentric coordinates
    lam = np.stack([1 - pts.sum(1), pts[:, 0], pts[:, 1], pts[:, 2]], axis=1)
    if order == 1:
        return np.broadcast_to(gl, (len(pts), 4, 3)).copy()
    out = np.zeros((len(pts), 10, 3))
    for a in range(4):
        out[:, a, :] = (4 * lam[:, a] - 1)[:, None] * gl[a]
    for e, (a, b) in enumerate(((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))):
        out[:, 4 + e, :] = 4 * (lam[:, a][:, None] * gl[b] + lam[:, b][:, None] * gl[a])
    return out


TET_QP_DEG2 = np.array([[0.1381966011250105] * 3, [0.5854101966249685, 0.1381966011250105, 0.1381966011250105],
                        [0.1381966011250105, 0.5854101966249685, 0.1381966011250105],
                        [0.1381966011250105, 0.1381966011250105, 0.5854101966249685]])
TET_QP_DEG1 = np.array([[0.25, 0.25, 0.25]])
