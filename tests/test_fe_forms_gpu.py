"""Fused flux/tangent -> element residual/stiffness contraction on the GPU against its oracle:
element vectors and matrices bit for bit, the device-side global assembly (fp64 atomics) to rounding, chained
u -> gradients -> constitutive update -> forms without the tangent ever visiting the host."""
import numpy as np
import pytest

from oracle import fe_forms as ff
from oracle import fe_gradient as fg

pytestmark = pytest.mark.gpu


def field(nodes, amp):
    x, y, z = nodes.T
    return amp * np.stack([x * y + 0.5 * z * z + 0.3 * x, -2 * y * z + 0.3 * x * x - 0.2 * y, 0.7 * x * z - 0.4 * y * y + 0.1 * z], axis=1)


def make_material(jm, finite):
    el = jm.LinearElasticIsotropic(E=70e3, nu=0.3)
    if finite:
        return jm.CUDAMaterial(jm.FeFpJ2Plasticity(elasticity=el, yield_stress=jm.VoceHardening(sig0=500.0, sigu=750.0, b=1000.0))), "PK1"
    return jm.CUDAMaterial(jm.vonMisesIsotropicHardening(elasticity=el, yield_stress=jm.VoceHardening(sig0=350.0, sigu=500.0, b=1e3))), "stress"


def device_outputs(mat, fname):
    flux = np.ascontiguousarray(mat.device_view(fname).cpu().numpy().T)
    ct = np.ascontiguousarray(mat.device_tangent().cpu().numpy().T)
    return flux, ct


@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("finite", [False, True])
def test_tet_element_forms_and_assembly(jm, order, finite):
    from dolfinx_materials_b200.fe import AssembledSystem, ElementForms, GradientEvaluator

    coords, gd, ud, nodes = fg.box_tets(7, 5, 6, order)  # cell count not a multiple of the cells per CTA
    qp = fg.TET_QP_DEG1 if order == 1 else fg.TET_QP_DEG2
    w = np.full(len(qp), 1.0 / 6.0 / len(qp))
    dphi = fg.tet_dphi(qp, order)
    nc, nqp = len(gd), len(qp)
    mat, fname = make_material(jm, finite)
    mat.set_data_manager(nc * nqp)
    ge = GradientEvaluator(mat, coords, gd, ud, dphi, tdim=3)
    forms = ElementForms(ge, w)
    with pytest.raises(Exception):
        forms.compute()  # no update has been run yet
    kind = 1 if finite else 0
    rowptr, colidx = ff.sparsity(ud, len(nodes), 3)
    system = AssembledSystem(forms, rowptr, colidx)
    for step, amp in enumerate((0.012, 0.03)):
        ge.eval(field(nodes, amp).ravel())
        stats = mat.integrate_resident()
        flux, ct = device_outputs(mat, fname)
        fe_ref, ke_ref = ff.element_forms(coords, gd, ud, dphi, w, flux, ct, kind, 3)
        fe, ke = forms.compute()
        assert np.array_equal(fe, fe_ref)
        assert np.array_equal(ke, ke_ref)
        fe_only, none = forms.compute(matrix=False)
        assert none is None and np.array_equal(fe_only, fe_ref)
        # global assembly, no constraints
        system.set_bc(None)
        system.assemble()
        vals, b = (x.copy() for x in system.get())
        b_ref, A_ref = ff.assemble(ud, fe_ref, ke_ref, len(nodes), 3)
        scale = np.abs(ke_ref).max()
        import scipy.sparse as sp

        A = sp.csr_matrix((vals, colidx, rowptr), shape=A_ref.shape)
        assert abs(A - A_ref).max() <= 1e-12 * scale
        assert np.allclose(b, b_ref, rtol=0, atol=1e-12 * np.abs(fe_ref).max())
        # with Dirichlet rows/columns (x = 0 face clamped)
        bc = np.repeat(nodes[:, 0] < 1e-9 + 0.03 * 1.0, 3)
        system.set_bc(bc)
        system.assemble()
        vals, b = (x.copy() for x in system.get())
        b_ref, A_ref = ff.assemble(ud, fe_ref, ke_ref, len(nodes), 3, bc=bc)
        A = sp.csr_matrix((vals, colidx, rowptr), shape=A_ref.shape)
        assert bc.any() and abs(A - A_ref).max() <= 1e-12 * scale
        assert np.all(b[bc] == 0) and np.allclose(b, b_ref, rtol=0, atol=1e-12 * np.abs(fe_ref).max())
        # residual only (what SNES asks for at a function evaluation): the thread-per-(cell, basis function) kernel;
        # the matrix values of the previous pass are left alone
        system.assemble(matrix=False)
        vals_kept, b_only = (x.copy() for x in system.get())
        assert np.all(b_only[bc] == 0) and np.allclose(b_only, b_ref, rtol=0, atol=1e-12 * np.abs(fe_ref).max())
        assert np.array_equal(vals_kept, vals)
        # inhomogeneous constraint: constrained columns move to the right-hand side (apply_lifting + set_bc)
        lift = np.where(bc, np.sin(np.arange(bc.size) * 0.37) * 1e-3, 0.0)
        system.set_lifting(lift)
        system.assemble()
        vals2, b = (x.copy() for x in system.get())
        b_ref, _ = ff.assemble(ud, fe_ref, ke_ref, len(nodes), 3, bc=bc, lift=lift)
        assert np.array_equal(vals2, vals) or np.allclose(vals2, vals, rtol=0, atol=1e-12 * scale)
        assert np.array_equal(b[bc], lift[bc]) and np.allclose(b, b_ref, rtol=0, atol=1e-12 * np.abs(b_ref).max())
        with pytest.raises(Exception, match="matrix pass"):
            system.assemble(matrix=False)
        system.set_lifting(None)
        # after update() the forms still read the flux of the converged step (s0 <- s1 swap)
        mat.data_manager.update()
        fe_after, _ = forms.compute(matrix=False)
        assert np.array_equal(fe_after, fe_ref)
    assert stats.n_plastic > 0


def test_triangles_generic_path_and_errors(jm):
    from dolfinx_materials_b200.fe import AssembledSystem, ElementForms, GradientEvaluator

    nx = 23
    xs = np.linspace(0, 1, nx + 1)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    c2 = np.stack([X.ravel() + 0.05 * Y.ravel() ** 2, Y.ravel(), np.zeros(X.size)], axis=1)
    nid = lambda i, j: i * (nx + 1) + j  # noqa: E731
    tri = np.array([[nid(i, j), nid(i + 1, j), nid(i + 1, j + 1)] for i in range(nx) for j in range(nx)]
                   + [[nid(i, j), nid(i + 1, j + 1), nid(i, j + 1)] for i in range(nx) for j in range(nx)], dtype=np.int32)
    dphi = np.broadcast_to(np.array([[-1.0, -1.0], [1, 0], [0, 1]]), (1, 3, 2)).copy()
    w = np.array([0.5])
    u = (0.02 * np.stack([c2[:, 0] * c2[:, 1], c2[:, 0] ** 2 - c2[:, 1]], axis=1)).ravel()
    mat, fname = make_material(jm, False)
    mat.set_data_manager(len(tri))
    ge = GradientEvaluator(mat, c2, tri, tri, dphi, tdim=2)
    with pytest.raises(ValueError):
        ElementForms(ge, np.array([0.25, 0.25]))
    forms = ElementForms(ge, w)
    ge.eval(u)
    stats = mat.integrate_resident()
    assert stats.n_plastic > 0
    flux, ct = device_outputs(mat, fname)
    fe_ref, ke_ref = ff.element_forms(c2, tri, tri, dphi, w, flux, ct, 0, 2)
    fe, ke = forms.compute()
    assert np.array_equal(fe, fe_ref) and np.array_equal(ke, ke_ref)
    # a pattern that lacks entries is reported, not silently dropped
    rowptr, colidx = ff.sparsity(tri[: len(tri) // 2], len(c2), 2)
    system = AssembledSystem(forms, rowptr, colidx)
    with pytest.raises(Exception, match="no slot"):
        system.assemble()

    # generic element path: P1 tets with a padded zero-gradient 5th dof
    coords, gd, ud, nodes = fg.box_tets(4, 3, 3, 1)
    ud5 = np.concatenate([ud, ud[:, :1]], axis=1)
    dphi5 = np.concatenate([fg.tet_dphi(fg.TET_QP_DEG1, 1), np.zeros((1, 1, 3))], axis=1)
    m3, f3 = make_material(jm, True)
    m3.set_data_manager(len(gd))
    g3 = GradientEvaluator(m3, coords, gd, ud5, dphi5, tdim=3)
    forms3 = ElementForms(g3, np.array([1.0 / 6.0]))
    g3.eval(field(nodes, 0.03).ravel())
    m3.integrate_resident()
    flux, ct = device_outputs(m3, f3)
    fe_ref, ke_ref = ff.element_forms(coords, gd, ud5, dphi5, np.array([1.0 / 6.0]), flux, ct, 1, 3)
    fe, ke = forms3.compute()
    assert np.array_equal(fe, fe_ref) and np.array_equal(ke, ke_ref)
