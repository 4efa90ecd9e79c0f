"""GPU gradient evaluation against its oracle, bit for bit, and chained with the
constitutive update: u (host) -> gradients (device) -> integrate_resident == oracle(integrate(oracle grads))."""
import numpy as np
import pytest

from oracle import fe_gradient as fg
from oracle import fefp
from oracle import small_strain as ss

pytestmark = pytest.mark.gpu


def field(nodes, amp):
    x, y, z = nodes.T
    return amp * np.stack([x * y + 0.5 * z * z + 0.3 * x, -2 * y * z + 0.3 * x * x - 0.2 * y, 0.7 * x * z - 0.4 * y * y + 0.1 * z], axis=1)


@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("finite", [False, True])
def test_tet_gradients_and_update(jm, order, finite):
    from dolfinx_materials_b200.fe import GradientEvaluator

    coords, gd, ud, nodes = fg.box_tets(7, 6, 5, order)
    qp = fg.TET_QP_DEG1 if order == 1 else fg.TET_QP_DEG2
    dphi = fg.tet_dphi(qp, order)
    n = len(gd) * len(qp)
    el = jm.LinearElasticIsotropic(E=70e3, nu=0.3)
    if finite:
        props = dict(E=70e3, nu=0.3, sig0=500.0, sigu=750.0, b=1000.0)
        mat = jm.CUDAMaterial(jm.FeFpJ2Plasticity(elasticity=el, yield_stress=jm.VoceHardening(sig0=500.0, sigu=750.0, b=1000.0)))
        gname, st, integ, flux = "F", fefp.virgin_state(n), fefp.integrate, "PK1"
    else:
        props = dict(E=70e3, nu=0.3, sig0=350.0, sigu=500.0, b=1e3)
        mat = jm.CUDAMaterial(jm.vonMisesIsotropicHardening(elasticity=el, yield_stress=jm.VoceHardening(sig0=350.0, sigu=500.0, b=1e3)))
        gname, st, integ, flux = "strain", ss.zero_state(n), ss.integrate, "stress"
    mat.set_data_manager(n)
    ge = GradientEvaluator(mat, coords, gd, ud, dphi, tdim=3)
    for amp in (0.01, 0.03):
        u = field(nodes, amp).ravel()
        ge.eval(u)
        g_ref = fg.evaluate(coords, gd, ud, u, dphi, 1 if finite else 0, 3)
        assert np.array_equal(mat.device_view(gname, gen=1).cpu().numpy().T, g_ref)
        stats = mat.integrate_resident()
        ref = integ(g_ref, st, props)
        assert np.array_equal(mat.device_view(flux).cpu().numpy().T, ref[flux])
        assert stats.n_plastic == int(ref["flag"].sum()) and stats.n_fail == 0
        mat.data_manager.update()
        st = {k: ref[k] for k in st}
    assert stats.n_plastic > 0


def test_generic_element_path_and_2d(jm):
    """ndofs_cell outside the compiled P1/P2 set takes the generic loop (here: P1 tets with a padded
    5th zero-gradient dof); triangles exercise tdim = 2 with plane-strain padding."""
    from dolfinx_materials_b200.fe import GradientEvaluator

    coords, gd, ud, nodes = fg.box_tets(4, 3, 3, 1)
    ud5 = np.concatenate([ud, ud[:, :1]], axis=1)
    dphi5 = np.concatenate([fg.tet_dphi(fg.TET_QP_DEG1, 1), np.zeros((1, 1, 3))], axis=1)
    n = len(gd)
    mat = jm.CUDAMaterial(jm.ElasticBehavior(elasticity=jm.LinearElasticIsotropic(E=70e3, nu=0.3)))
    mat.set_data_manager(n)
    u = field(nodes, 0.02).ravel()
    GradientEvaluator(mat, coords, gd, ud5, dphi5, tdim=3).eval(u)
    assert np.array_equal(mat.device_view("strain", gen=1).cpu().numpy().T, fg.evaluate(coords, gd, ud5, u, dphi5, 0, 3))
    # 2-D: structured triangles of the unit square
    nx = 40
    xs = np.linspace(0, 1, nx + 1)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    c2 = np.stack([X.ravel(), Y.ravel(), np.zeros(X.size)], axis=1)
    nid = lambda i, j: i * (nx + 1) + j  # noqa: E731
    tri = np.array([[nid(i, j), nid(i + 1, j), nid(i + 1, j + 1)] for i in range(nx) for j in range(nx)]
                   + [[nid(i, j), nid(i + 1, j + 1), nid(i, j + 1)] for i in range(nx) for j in range(nx)], dtype=np.int32)
    dphi = np.broadcast_to(np.array([[-1.0, -1.0], [1, 0], [0, 1]]), (1, 3, 2)).copy()
    u2 = (0.01 * np.stack([c2[:, 0] * c2[:, 1], c2[:, 0] ** 2 - c2[:, 1]], axis=1)).ravel()
    m2 = jm.CUDAMaterial(jm.ElasticBehavior(elasticity=jm.LinearElasticIsotropic(E=70e3, nu=0.3)))
    m2.set_data_manager(len(tri))
    GradientEvaluator(m2, c2, tri, tri, dphi, tdim=2).eval(u2)
    ref = fg.evaluate(c2, tri, tri, u2, dphi, 0, 2)
    got = m2.device_view("strain", gen=1).cpu().numpy().T
    assert np.array_equal(got, ref) and np.count_nonzero(got[:, [2, 4, 5]]) == 0
    with pytest.raises(Exception):
        GradientEvaluator(mat, c2, tri, tri, dphi, tdim=2).eval(u2)  # point count mismatch is an error
