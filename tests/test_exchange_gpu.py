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


@pytest.mark.parametrize("fefp", [False, True])
def test_pipelined_subset_exchange_equals_reference_sequence(jm, monkeypatch, fefp):
    """A subset map large enough for the pipelined path (chunks of whole cells through integrate_range_into, host
    gather / scatter of the neighbouring chunks overlapped on a helper thread): same arrays as the reference sequence.
    PIPELINE_POINTS is lowered so that a small mesh already has many (ragged) chunks."""
    from dolfinx_materials_b200.exchange import QuadratureExchange

    monkeypatch.setattr(QuadratureExchange, "PIPELINE_POINTS", 3000)
    ncell, nqp = 9001, 4
    ntot = ncell * nqp
    cells = np.sort(np.random.default_rng(3).choice(ncell, 6100, replace=False))
    gname, gdim = ("F", 9) if fefp else ("strain", 6)
    g0 = np.tile([1, 1, 1, 0, 0, 0, 0, 0, 0.0], (ntot, 1)) if fefp else np.zeros((ntot, 6))
    gen = (lambda k: synth.defgrad(ntot, 1, 3e-2, k, 2)) if fefp else (lambda k: synth.strain(ntot, 1, 1.25e-2, k, 2))
    ref = QuadratureMapReplay(ncell, nqp, material(jm, fefp), cells=cells)
    ref.register_gradient(gname, g0)
    mat = material(jm, fefp)
    grad, flux, jac = g0.copy().ravel(), np.zeros(ntot * gdim), np.zeros(ntot * gdim * gdim)
    isv = {k: np.zeros(ntot * d) for k, d in mat.internal_state_variables.items()}
    ex = QuadratureExchange(mat, ncell, nqp, {gname: grad}, {mat.flux_names[0]: flux}, isv, jac, cells=cells)
    assert ex._chunks is not None and len(ex._chunks) > 8
    if fefp:
        ref.update_initial_state("be_bar", np.array([1, 1, 1, 0, 0, 0.0]))
        ex.update_initial_state("be_bar", np.array([1, 1, 1, 0, 0, 0.0]))
    ref.update()
    ex.update()
    for step in (1, 2):
        g = gen(step)
        ref.set_gradient_values(gname, g)
        grad[:] = g.ravel()
        ref.update()
        stats = ex.update()
        assert np.array_equal(flux, ref.fluxes[mat.flux_names[0]].array)
        assert np.array_equal(jac, ref.jacobian_flatten.array)
        assert stats.n_fail == 0 and stats.n_points == len(cells) * nqp
        ref.advance()
        ex.advance()
        assert np.array_equal(flux, ref.fluxes[mat.flux_names[0]].array)
        for k in isv:
            assert np.array_equal(isv[k], ref.internal_state_variables[k].array)
    assert stats.n_plastic > 0
    ex.close()


def test_integrate_range_covers_the_batch_like_one_call(jm):
    """dxm_integrate_range: three ranges (even starts, ragged last) == one integrate over the whole handle, with the
    statistics adding up; odd starts and overlong ranges are rejected."""
    from dolfinx_materials_b200._lib import DxmError

    n = 100_001
    whole, parts = material(jm), material(jm)
    whole.set_data_manager(n)
    parts.set_data_manager(n)
    for k in (1, 2):
        eps = synth.strain(n, 2, 1.25e-2, k, 2)
        f0, i0, c0 = [x.copy() for x in whole.integrate(eps)]
        f1, i1, c1 = np.empty((n, 6)), np.empty((n, 7)), np.empty((n, 36))
        n_pl = 0
        for a, b in ((0, 40_000), (40_000, 70_002), (70_002, n)):
            st = parts.integrate_range_into(a, b - a, eps[a:b], f1[a:b], i1[a:b], c1[a:b])
            n_pl += st.n_plastic
            assert st.n_points == b - a
        assert np.array_equal(f0, f1) and np.array_equal(i0, i1) and np.array_equal(c0.reshape(n, 36), c1)
        assert n_pl == whole.last_stats.n_plastic
        whole.data_manager.update()
        parts.data_manager.update()
    with pytest.raises(DxmError):
        parts.integrate_range_into(1, 10, eps[1:11], f1[1:11])
    with pytest.raises(DxmError):
        parts.integrate_range_into(n - 5, 10, eps[:10], f1[:10])
