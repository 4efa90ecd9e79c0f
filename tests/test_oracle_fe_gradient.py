"""The FE gradient-evaluation oracle reproduces exact gradients (CPU only): affine fields on P1 and P2 tets,
quadratic fields on P2 tets (evaluated at the physical quadrature points), for both gradient kinds."""
import numpy as np

from oracle import fe_gradient as fg

R2 = 2 ** -0.5


def phys_points(coords, gd, qp):
    x = coords[gd]  # (nc, 4, 3)
    lam = np.stack([1 - qp.sum(1), qp[:, 0], qp[:, 1], qp[:, 2]], axis=1)  # (nqp, 4)
    return np.einsum("qv,cvi->cqi", lam, x).reshape(-1, 3)


def test_affine_field_p1_and_p2():
    A = np.array([[0.01, -0.02, 0.005], [0.03, 0.0, -0.01], [0.002, 0.004, -0.015]])
    for order, qp in ((1, fg.TET_QP_DEG1), (2, fg.TET_QP_DEG2)):
        coords, gd, ud, nodes = fg.box_tets(3, 2, 2, order)
        u = (nodes @ A.T + np.array([0.1, 0.2, 0.3])).ravel()
        dphi = fg.tet_dphi(qp, order)
        F = fg.evaluate(coords, gd, ud, u, dphi, 1, 3)
        G = A
        expect = np.array([1 + G[0, 0], 1 + G[1, 1], 1 + G[2, 2], G[0, 1], G[1, 0], G[0, 2], G[2, 0], G[1, 2], G[2, 1]])
        assert F.shape == (len(gd) * len(qp), 9) and np.abs(F - expect).max() < 1e-13
        eps = fg.evaluate(coords, gd, ud, u, dphi, 0, 3)
        e = 0.5 * (G + G.T)
        expect = np.array([e[0, 0], e[1, 1], e[2, 2], 2 * R2 * e[0, 1], 2 * R2 * e[0, 2], 2 * R2 * e[1, 2]])
        assert np.abs(eps - expect).max() < 1e-13


def test_quadratic_field_p2_exact():
    coords, gd, ud, nodes = fg.box_tets(2, 3, 2, 2)
    x, y, z = nodes.T
    u = np.stack([0.1 * x * y + 0.05 * z * z, -0.2 * y * z + 0.03 * x * x, 0.07 * x * z - 0.04 * y * y], axis=1).ravel()
    F = fg.evaluate(coords, gd, ud, u, fg.tet_dphi(fg.TET_QP_DEG2, 2), 1, 3)
    xq, yq, zq = phys_points(coords, gd, fg.TET_QP_DEG2).T
    G = np.zeros((len(xq), 3, 3))
    G[:, 0] = np.stack([0.1 * yq, 0.1 * xq, 0.1 * zq], axis=1)
    G[:, 1] = np.stack([0.06 * xq, -0.2 * zq, -0.2 * yq], axis=1)
    G[:, 2] = np.stack([0.07 * zq, -0.08 * yq, 0.07 * xq], axis=1)
    expect = np.stack([1 + G[:, 0, 0], 1 + G[:, 1, 1], 1 + G[:, 2, 2], G[:, 0, 1], G[:, 1, 0], G[:, 0, 2], G[:, 2, 0],
                       G[:, 1, 2], G[:, 2, 1]], axis=1)
    assert np.abs(F - expect).max() < 1e-12


def test_plane_strain_padding_2d():
    # two triangles of the unit square, P1, affine field: strain = [exx, eyy, 0, r2*exy, 0, 0]
    coords = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0.0]])
    gd = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.int32)
    A = np.array([[0.01, 0.004], [-0.002, 0.03]])
    u = (coords[:, :2] @ A.T).ravel()
    dphi = np.broadcast_to(np.array([[-1.0, -1.0], [1, 0], [0, 1]]), (1, 3, 2)).copy()
    eps = fg.evaluate(coords, gd, gd, u, dphi, 0, 2)
    expect = np.array([A[0, 0], A[1, 1], 0.0, (A[0, 1] + A[1, 0]) * R2, 0.0, 0.0])
    assert np.abs(eps - expect).max() < 1e-15
    F = fg.evaluate(coords, gd, gd, u, dphi, 1, 2)
    assert np.abs(F - np.array([1 + A[0, 0], 1 + A[1, 1], 1.0, A[0, 1], A[1, 0], 0, 0, 0, 0])).max() < 1e-15
