"""QuadratureExchange (the GPU-aware update/advance of SURVEY 8(f) rank 1) leaves exactly the same values in
the same Function arrays as the reference's call sequence replayed by tests/qmap_replay.py."""
import numpy as np
import pytest

from oracle import synth
from qmap_replay import QuadratureMapReplay

pytestmark = pytest.mark.gpu
VOCE = dict(E=70e3, nu=0.3, sig0=350.0, sigu=500.0, b=1e3)


def material(jm, fefp=False):
    el = jm.LinearElasticIsotropic(E=70e3, nu=0.3)
    if fefp:
        return jm.CUDAMaterial(jm.FeFpJ2Plasticity(elasticity=el, yield_stress=jm.VoceHardening(sig0=500.0, sigu=750.0, b=1000.0)))
    return jm.CUDAMaterial(jm.vonMisesIsotropicHardening(elasticity=el, yield_stress=jm.VoceHardening(sig0=350.0, sigu=500.0, b=1e3)))


@pytest.mark.parametrize("subset", [False, True])
@pytest.mark.parametrize("fefp", [False, True])
def test_exchange_equals_reference_sequence(jm, subset, fefp):
    from dolfinx_materials_b200.exchange import QuadratureExchange

    ncell, nqp = 1200, 4
    ntot = ncell * nqp
    cells = np.arange(ncell)[::3] if subset else None
    gname, gdim = ("F", 9) if fefp else ("strain", 6)
    g0 = np.tile([1, 1, 1, 0, 0, 0, 0, 0, 0.0], (ntot, 1)) if fefp else np.zeros((ntot, 6))
    gen = (lambda k: synth.defgrad(ntot, 1, 3e-2, k, 3)) if fefp else (lambda k: synth.strain(ntot, 1, 1.25e-2, k, 3))

    ref = QuadratureMapReplay(ncell, nqp, material(jm, fefp), cells=cells)
    ref.register_gradient(gname, g0)

    mat = material(jm, fefp)
    grad = g0.copy().ravel()
    flux = np.zeros(ntot * gdim)
    isv = {k: np.zeros(ntot * d) for k, d in mat.internal_state_variables.items()}
    jac = np.zeros(ntot * gdim * gdim)
    ex = QuadratureExchange(mat, ncell, nqp, {gname: grad}, {mat.flux_names[0]: flux}, isv, jac, cells=cells)
    if fefp:
        ref.update_initial_state("be_bar", np.array([1, 1, 1, 0, 0, 0.0]))
        ex.update_initial_state("be_bar", np.array([1, 1, 1, 0, 0, 0.0]))
    ref.update()
    ex.update()
    for step in range(1, 4):
        for scale in (0.8, 1.0):
            g = g0 + scale * (gen(step) - g0)
            ref.set_gradient_values(gname, g)
            grad[:] = g.ravel()
            ref.update()
            stats = ex.update()
            assert np.array_equal(flux, ref.fluxes[mat.flux_names[0]].array)
            assert np.array_equal(jac, ref.jacobian_flatten.array)
            assert stats.n_fail == 0
        ref.advance()
        ex.advance()
        assert np.array_equal(flux, ref.fluxes[mat.flux_names[0]].array)
        for k in isv:
            assert np.array_equal(isv[k], ref.internal_state_variables[k].array)
    assert stats.n_plastic > 0
    ex.close()


def test_host_path_many_chunks_ragged(jm):
    """Host path with n > pipeline chunk (2^19) and a ragged last chunk, pageable and page-locked buffers."""
    from dolfinx_materials_b200.material import pin_array
    from oracle import small_strain as ss

    n = 2 * (1 << 19) + 12345
    mat = material(jm)
    mat.set_data_manager(n)
    eps = synth.strain(n, 2, 1.25e-2, 1, 1)
    ref = ss.integrate(eps, ss.zero_state(n), VOCE)
    flux, isv, Ct = mat.integrate(eps)  # pageable input, pinned outputs
    assert np.array_equal(flux, ref["stress"]) and np.array_equal(Ct, ref["Ct"])
    assert np.array_equal(isv[:, 0], ref["p"]) and np.array_equal(isv[:, 1:], ref["epsp"])
    f2, c2 = np.zeros((n, 6)), np.zeros((n, 36))  # pageable outputs
    mat.integrate_into(eps, f2, None, c2)
    assert np.array_equal(f2, ref["stress"]) and np.array_equal(c2, ref["Ct"].reshape(n, 36))
    unpin = pin_array(f2)
    f2[:] = 0
    mat.integrate_into(eps, f2, None, None)
    assert np.array_equal(f2, ref["stress"])
    unpin()


def test_exchange_strict_failure_behaviour(jm):
    """Non-finite gradients: strict mode raises AssertionError like the reference's NaN asserts
    (quadrature_map.py:322-324); non-strict warns (mfront.py:269-272 convention)."""
    from dolfinx_materials_b200.exchange import QuadratureExchange

    ncell, nqp = 50, 1
    for strict in (True, False):
        mat = material(jm)
        grad = synth.strain(ncell, 0, 1e-2, 1, 1).ravel()
        ex = QuadratureExchange(mat, ncell, nqp, {"strain": grad}, {"stress": np.zeros(ncell * 6)},
                                {"p": np.zeros(ncell), "epsp": np.zeros(ncell * 6)}, np.zeros(ncell * 36), strict=strict)
        ex.initialize_state()
        grad[7] = np.nan  # this iteration's evaluated gradients hold one non-finite entry (point 1)
        if strict:
            with pytest.raises(AssertionError):
                ex.update()
        else:
            with pytest.warns(jm.PerformanceWarning):
                assert ex.update().n_fail == 1
        ex.close()
