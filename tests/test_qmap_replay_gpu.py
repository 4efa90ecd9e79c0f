"""CUDAMaterial driven exactly as the reference's QuadratureMap drives a material (tests/qmap_replay.py):
Newton-like repeated update() from the same s0, advance() after "convergence", cell subsets, two maps on
disjoint halves of one mesh (the pattern of tests/mfront/test_multimaterials.py:97-105, :163-172)."""
import numpy as np
import pytest

from oracle import fefp, synth
from oracle import small_strain as ss
from qmap_replay import QuadratureMapReplay, _get_vals

pytestmark = pytest.mark.gpu
VOCE = dict(E=70e3, nu=0.3, sig0=350.0, sigu=500.0, b=1e3)


def voce_material(jm):
    return jm.CUDAMaterial(jm.vonMisesIsotropicHardening(
        elasticity=jm.LinearElasticIsotropic(E=VOCE["E"], nu=VOCE["nu"]),
        yield_stress=jm.VoceHardening(sig0=VOCE["sig0"], sigu=VOCE["sigu"], b=VOCE["b"])))


def test_newton_like_load_stepping_full_mesh(jm):
    ncell, nqp = 500, 4  # P2 tets: 4 points per cell (finite_strain_elastoplasticity.py:115-117)
    n = ncell * nqp
    qmap = QuadratureMapReplay(ncell, nqp, voce_material(jm))
    qmap.register_gradient("strain", np.zeros((n, 6)))
    # first update at u = 0, as the demos do before the solve (finite_strain_elastoplasticity.py:185):
    # initialize_state() captures the gradients evaluated at that moment into s0 (quadrature_map.py:281-295)
    qmap.update()
    assert np.count_nonzero(_get_vals(qmap.fluxes["stress"])) == 0
    st = ss.zero_state(n)
    for step in range(1, 4):
        # three "Newton iterations": perturbed gradients, always integrating from the same s0
        for it, scale in enumerate([0.7, 0.95, 1.0]):
            eps = scale * synth.strain(n, 0, 1.25e-2, step, 3)
            qmap.set_gradient_values("strain", eps)
            qmap.update()
            ref = ss.integrate(eps, st, VOCE)
            assert np.array_equal(_get_vals(qmap.fluxes["stress"]), ref["stress"])
            assert np.array_equal(qmap.jacobian_flatten.array.reshape(n, 36), ref["Ct"].reshape(n, 36))
            assert np.array_equal(_get_vals(qmap.internal_state_variables["p"])[:, 0], ref["p"])
            assert np.array_equal(_get_vals(qmap.internal_state_variables["epsp"]), ref["epsp"])
        qmap.advance()
        st = ss.advance(ref)
        assert np.array_equal(_get_vals(qmap.internal_state_variables["p"])[:, 0], st["p"])
    assert ref["flag"].mean() > 0.3


def test_two_maps_on_disjoint_cell_subsets_equal_one_map(jm):
    ncell, nqp = 301, 3
    n = ncell * nqp
    eps_all = synth.strain(n, 5, 1.25e-2, 1, 1)
    mono = QuadratureMapReplay(ncell, nqp, voce_material(jm))
    mono.register_gradient("strain", eps_all)
    mono.update()
    cells = np.arange(ncell)
    left, right = cells[cells % 2 == 0], cells[cells % 2 == 1]  # interleaved subsets -> real gather/scatter
    maps = []
    for sub in (left, right):
        q = QuadratureMapReplay(ncell, nqp, voce_material(jm), cells=sub)
        q.register_gradient("strain", eps_all)
        q.update()
        maps.append(q)
    sig_l, sig_r = (_get_vals(q.fluxes["stress"]) for q in maps)
    assert np.array_equal(sig_l + sig_r, _get_vals(mono.fluxes["stress"]))
    assert np.count_nonzero(sig_l * sig_r) == 0  # disjoint supports (test_multimaterials.py:171-172)
    assert np.array_equal(maps[0].jacobian_flatten.array + maps[1].jacobian_flatten.array, mono.jacobian_flatten.array)


def test_fefp_map_with_initial_state_update(jm):
    ncell, nqp = 200, 4
    n = ncell * nqp
    props = dict(E=70e3, nu=0.3, sig0=500.0, sigu=750.0, b=1000.0)
    mat = jm.CUDAMaterial(jm.FeFpJ2Plasticity(
        elasticity=jm.LinearElasticIsotropic(E=props["E"], nu=props["nu"]),
        yield_stress=jm.VoceHardening(sig0=props["sig0"], sigu=props["sigu"], b=props["b"])))
    qmap = QuadratureMapReplay(ncell, nqp, mat)
    ident = np.tile([1, 1, 1, 0, 0, 0, 0, 0, 0.0], (n, 1))
    qmap.register_gradient("F", ident)
    # the demo initialises be_bar explicitly (finite_strain_elastoplasticity.py:181); without it
    # initialize_state() would push the zero-initialised Function into s0
    qmap.update_initial_state("be_bar", np.array([1, 1, 1, 0, 0, 0.0]))
    qmap.update()  # "enforce compilation" call of the demo at F = I (finite_strain_elastoplasticity.py:185)
    assert np.abs(_get_vals(qmap.fluxes["PK1"])).max() == 0
    st = fefp.virgin_state(n)
    for step in range(1, 4):
        F = synth.defgrad(n, 2, 3e-2, step, 3)
        qmap.set_gradient_values("F", F)
        qmap.update()
        ref = fefp.integrate(F, st, props)
        assert np.array_equal(_get_vals(qmap.fluxes["PK1"]), ref["PK1"])
        assert np.array_equal(qmap.jacobian_flatten.array.reshape(n, 81), ref["Ct"].reshape(n, 81))
        qmap.advance()
        st = fefp.advance(ref)
        assert np.array_equal(_get_vals(qmap.internal_state_variables["be_bar"]), st["be_bar"])
    assert ref["flag"].mean() > 0.2


def test_multimaterial_hosford_matrix_and_voce_inclusions(jm):
    """The reference's multi-material set-up (demos/multimaterials/multimaterials.py:245-273): the matrix cells carry the
    Hosford (a = 10) + linear hardening law that the demo takes from MFront, the inclusion cells J2 + Voce, each
    through its own map on a cell subset; load stepping with advance().  Stresses and tangents of both maps equal the
    oracles on their own points, and add up to a field with disjoint supports."""
    from oracle import hosford as ho

    ncell, nqp = 400, 1  # P1 triangles, one point per cell (multimaterials.py:264)
    matrix_props = dict(E=70e3, nu=0.3, sig0=200.0, H=10.0, a=10)  # multimaterials.py:245-254
    incl_props = dict(E=90e3, nu=0.25, sig0=200.0, sigu=300.0, b=10.0)  # multimaterials.py:256-261
    cells = np.arange(ncell)
    incl = cells[(cells // 20) % 3 == 1]  # blocks of inclusion cells
    matrix = np.setdiff1d(cells, incl)
    m1 = jm.CUDAMaterial(jm.GeneralIsotropicHardening(
        elasticity=jm.LinearElasticIsotropic(E=matrix_props["E"], nu=matrix_props["nu"]),
        yield_stress=jm.LinearHardening(sig0=matrix_props["sig0"], H=matrix_props["H"]),
        equivalent_stress=jm.Hosford(a=10)))
    m2 = jm.CUDAMaterial(jm.vonMisesIsotropicHardening(
        elasticity=jm.LinearElasticIsotropic(E=incl_props["E"], nu=incl_props["nu"]),
        yield_stress=jm.VoceHardening(sig0=incl_props["sig0"], sigu=incl_props["sigu"], b=incl_props["b"])))
    q1 = QuadratureMapReplay(ncell, nqp, m1, cells=matrix)
    q2 = QuadratureMapReplay(ncell, nqp, m2, cells=incl)
    zero = np.zeros((ncell * nqp, 6))
    for q in (q1, q2):
        q.register_gradient("strain", zero)
        q.update()
    st1, st2 = ss.zero_state(len(matrix)), ss.zero_state(len(incl))
    for step in range(1, 4):
        eps = synth.strain(ncell * nqp, 8, 1.25e-2, step, 3)
        for q in (q1, q2):
            q.set_gradient_values("strain", eps)
            q.update()
        r1 = ho.integrate(eps[matrix], st1, matrix_props)
        r2 = ss.integrate(eps[incl], st2, incl_props)
        s1, s2 = _get_vals(q1.fluxes["stress"]), _get_vals(q2.fluxes["stress"])
        assert np.array_equal(s1[matrix], r1["stress"]) and np.array_equal(s2[incl], r2["stress"])
        assert np.count_nonzero(s1[incl]) == 0 and np.count_nonzero(s2[matrix]) == 0
        assert np.array_equal(q1.jacobian_flatten.array.reshape(-1, 36)[matrix], r1["Ct"].reshape(-1, 36))
        assert np.array_equal(q2.jacobian_flatten.array.reshape(-1, 36)[incl], r2["Ct"].reshape(-1, 36))
        for q in (q1, q2):
            q.advance()
        st1, st2 = ss.advance(r1), ss.advance(r2)
        assert np.array_equal(_get_vals(q1.internal_state_variables["p"])[matrix, 0], st1["p"])
    assert r1["flag"].mean() > 0.3 and r2["flag"].mean() > 0.3
