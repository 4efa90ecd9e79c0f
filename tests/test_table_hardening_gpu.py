"""GPU parity of the tabulated (piecewise-linear) isotropic hardening behaviour (DXM_J2_TABLE)
against the oracle: bit-identical flags, segment-crossing counts, stress, state and tangent; per-point elastic
properties; the callable-sampling front end; error paths."""
import numpy as np
import pytest

from oracle import small_strain as ss
from oracle import synth

pytestmark = pytest.mark.gpu

PK = np.array([0.0, 2e-4, 1e-3, 4e-3, 2e-2, 0.2])
SK = np.array([300.0, 340.0, 390.0, 430.0, 470.0, 520.0])


def check_history(m, props, n, amp, K):
    st = ss.zero_state(n)
    for k in range(1, K + 1):
        eps = synth.strain(n, 0, amp, k, K)
        flux, isv, Ct = m.integrate(eps)
        ref = ss.integrate(eps, st, props)
        flag, n_iter, resid, fail = m.diagnostics()
        assert np.array_equal(flag, ref["flag"]) and np.array_equal(n_iter, ref["n_iter"]) and np.array_equal(fail, ref["fail"])
        assert np.array_equal(flux, ref["stress"]) and np.array_equal(Ct, ref["Ct"])
        assert np.array_equal(isv[:, 0], ref["p"]) and np.array_equal(isv[:, 1:], ref["epsp"])
        assert m.last_stats.n_plastic == int(ref["flag"].sum()) and m.last_stats.max_iter == int(ref["n_iter"].max())
        m.data_manager.update()
        st = ss.advance(ref)
    return ref


def test_table_hardening_matches_oracle(jm):
    n = 20_011
    beh = jm.vonMisesIsotropicHardening(elasticity=jm.LinearElasticIsotropic(E=70e3, nu=0.3),
                                        yield_stress=jm.TabulatedHardening(p=PK, sig=SK))
    m = jm.CUDAMaterial(beh)
    m.set_data_manager(n)
    m.enable_diagnostics()
    ref = check_history(m, dict(E=70e3, nu=0.3, table=(PK, SK)), n, 1.25e-2, 4)
    assert 0.4 < ref["flag"].mean() < 0.95 and ref["n_iter"].max() >= 2
    # resident path + packed tangent view agree with the host path
    m2 = jm.CUDAMaterial(beh)
    m2.set_data_manager(n)
    eps = synth.strain(n, 0, 1.25e-2, 1, 1)
    m2.gradient_buffer().copy_(__import__("torch").from_numpy(np.ascontiguousarray(eps.T)))
    m2.integrate_resident()
    r1 = ss.integrate(eps, ss.zero_state(n), dict(E=70e3, nu=0.3, table=(PK, SK)))
    assert np.array_equal(m2.device_tangent().cpu().numpy().T.reshape(n, 6, 6), r1["Ct"])


def test_table_with_per_point_elasticity_and_callable_front_end(jm):
    n = 5_003
    E = np.linspace(60e3, 90e3, n)
    voce = lambda p: 350.0 + 150.0 * (1 - np.exp(-1e3 * p))  # noqa: E731
    tab = jm.TabulatedHardening.from_callable(voce, p_max=0.2, n=64)
    assert len(tab.p) == 64 and tab.p[0] == 0.0 and tab.sig[0] == 350.0
    m = jm.CUDAMaterial(jm.vonMisesIsotropicHardening(elasticity=jm.LinearElasticIsotropic(E=70e3, nu=0.3), yield_stress=tab))
    m.set_data_manager(n)
    m.enable_diagnostics()
    m.update_material_property("E", E)
    ref = check_history(m, dict(E=E, nu=0.3, table=(tab.p, tab.sig)), n, 1.25e-2, 3)
    # ... and the sampled law is close to the Voce behaviour it came from
    mv = jm.CUDAMaterial(jm.vonMisesIsotropicHardening(elasticity=jm.LinearElasticIsotropic(E=70e3, nu=0.3),
                                                       yield_stress=jm.VoceHardening(sig0=350.0, sigu=500.0, b=1e3)))
    mv.set_data_manager(n)
    mv.update_material_property("E", E)
    for k in range(1, 4):
        fv, _, _ = mv.integrate(synth.strain(n, 0, 1.25e-2, k, 3))
        mv.data_manager.update()
    assert np.abs(fv - ref["stress"]).max() < 3e-4 * np.abs(fv).max()


def test_table_error_paths(jm):
    from dolfinx_materials_b200._lib import DxmError

    el = jm.LinearElasticIsotropic(E=70e3, nu=0.3)
    for p, s in (([0.0], [1.0]), ([0.1, 0.2], [1.0, 2.0]), ([0.0, 0.2, 0.1], [1.0, 2.0, 3.0]), (np.linspace(0, 1, 65), np.ones(65))):
        m = jm.CUDAMaterial(jm.vonMisesIsotropicHardening(elasticity=el, yield_stress=jm.TabulatedHardening(p=p, sig=s)))
        with pytest.raises(DxmError):
            m.set_data_manager(8)
