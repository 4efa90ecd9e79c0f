"""Per-call statistics folded and published by the update kernel itself (no memset / copy / stream synchronisation on
the host side): they must equal what the per-point diagnostics say, call after call, on every path that launches the
kernel -- resident, host (chunked pipeline: several launches, the last one publishes), ranges, asynchronous calls."""
import numpy as np
import pytest

from oracle import synth

pytestmark = pytest.mark.gpu


def _material(jm, kind):
    el = jm.LinearElasticIsotropic(E=70e3, nu=0.3)
    if kind == "j2":
        return jm.CUDAMaterial(jm.vonMisesIsotropicHardening(elasticity=el, yield_stress=jm.VoceHardening(sig0=350.0, sigu=500.0, b=1e3)))
    if kind == "hosford":
        return jm.CUDAMaterial(jm.GeneralIsotropicHardening(elasticity=el, yield_stress=jm.LinearHardening(sig0=200.0, H=10.0),
                                                            equivalent_stress=jm.Hosford(a=10)))
    return jm.CUDAMaterial(jm.FeFpJ2Plasticity(elasticity=el, yield_stress=jm.VoceHardening(sig0=500.0, sigu=750.0, b=1000.0)))


def _check(m, s, n):
    flag, n_iter, resid, fail = m.diagnostics()
    assert s.n_points == n
    assert s.n_plastic == int(flag.sum())
    assert s.n_fail == int(fail.sum())
    assert s.max_iter == int(n_iter.max())
    assert s.max_residual == float(resid.max())


@pytest.mark.parametrize("kind", ["j2", "hosford", "fefp"])
@pytest.mark.parametrize("n", [1, 37, 5000, 700_001])
def test_published_stats_equal_diagnostics(jm, kind, n):
    m = _material(jm, kind)
    m.set_data_manager(n)
    m.enable_diagnostics()
    K = 3
    for k in range(1, K + 1):
        g = synth.defgrad(n, 0, 3e-2, k, K) if kind == "fefp" else synth.strain(n, 0, 1.25e-2, k, K)
        m.integrate(g)  # host path: chunked pipeline for the large batch
        _check(m, m.last_stats, n)
        s = m.integrate_resident()  # same gradients, resident path
        _check(m, s, n)
        assert m.integrate_resident(wait=False) is None  # asynchronous: published record fetched later
        _check(m, m.fetch_stats(), n)
        m.data_manager.update()
    assert m.last_stats.n_plastic > 0 or n < 10


def test_stats_of_ranges_and_failures(jm):
    n = 200_000
    m = _material(jm, "j2")
    m.set_data_manager(n)
    m.enable_diagnostics()
    eps = synth.strain(n, 0, 1.25e-2, 1, 1)
    eps[1234, 0] = np.nan  # one failed point
    flux, isv, ct = np.empty((n, 6)), np.empty((n, 7)), np.empty((n, 36))
    tot_plastic = tot_fail = 0
    for lo, hi in ((0, 65536), (65536, 131072), (131072, n)):
        s = m.integrate_range_into(lo, hi - lo, eps[lo:hi], flux[lo:hi], isv[lo:hi], ct[lo:hi])
        assert s.n_points == hi - lo
        tot_plastic += s.n_plastic
        tot_fail += s.n_fail
    flag, _, _, fail = m.diagnostics()
    assert tot_plastic == int(flag.sum()) and tot_fail == int(fail.sum()) == 1
    with pytest.warns(jm.PerformanceWarning):
        m.integrate(eps)
    assert m.last_stats.n_fail == 1


def test_timing_switch(jm):
    m = _material(jm, "j2")
    m.set_data_manager(1000)
    m.synth_gradients(0, 1e-2, 1, 1)
    assert m.integrate_resident().kernel_ms == 0.0  # small batch: no events by default
    m.enable_timing(1)
    assert m.integrate_resident().kernel_ms > 0.0
    m.enable_timing(0)
    assert m.integrate_resident().kernel_ms == 0.0
