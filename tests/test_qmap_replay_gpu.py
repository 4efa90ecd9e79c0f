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
