"""GPU parity of the finite-strain FeFp kernel against the CPU oracle through the public API; the
first test has the shape of the reference's own tests/test_FeFp_jax.py."""
import numpy as np
import pytest

from oracle import fefp, synth

pytestmark = pytest.mark.gpu
RTOL = 1e-10
PROPS = dict(E=70e3, nu=0.3, sig0=500.0, sigu=750.0, b=1000.0)


def make(jm, n, props=PROPS, diag=True):
    elastic_model = jm.LinearElasticIsotropic(E=props["E"], nu=props["nu"])
    behavior = jm.FeFpJ2Plasticity(
        elasticity=elastic_model, yield_stress=jm.VoceHardening(sig0=props["sig0"], sigu=props["sigu"], b=props["b"])
    )
    material = jm.CUDAMaterial(behavior)
    material.set_data_manager(n)
    if diag:
        material.enable_diagnostics()
    return material


def compare(material, F, st):
    P, isv, Ct = material.integrate(F, 0)
    ref = fefp.integrate(F, st, PROPS)
    flag, n_iter, resid, fail = material.diagnostics()
    assert np.array_equal(flag, ref["flag"])
    assert np.array_equal(n_iter, ref["n_iter"])
    assert np.array_equal(fail, ref["fail"])
    assert np.array_equal(P, ref["PK1"])
    assert np.array_equal(isv[:, 0], ref["p"]) and np.array_equal(isv[:, 1:], ref["be_bar"])
    assert np.array_equal(Ct, ref["Ct"])
    assert np.array_equal(resid, ref["resid"])
    np.testing.assert_allclose(P, ref["PK1"], rtol=RTOL, atol=0)
    return ref


def test_FeFp_plasticity(jm, Nbatch=10):
    """Same script as the reference's tests/test_FeFp_jax.py:6-33 (which asserts nothing) -- here every
    step is compared with the oracle and the end state with the independently obtained values."""
    material = make(jm, Nbatch)
    eps = 2e-2
    Nsteps = 20
    st = fefp.virgin_state(Nbatch)
    for t in np.linspace(0, 1.0, Nsteps)[1:]:
        F = np.zeros((Nbatch, 9))
        F[:, 0] = 1 + eps * t
        F[:, [1, 2]] = 1 - eps / 2 * t
        ref = compare(material, F, st)
        material.data_manager.update()
        st = fefp.advance(ref)
    state = material.get_final_state_dict()
    assert abs(state["p"][0, 0] - 1.076097e-2) < 5e-9 and abs(state["PK1"][0, 0] - 473.1527) < 5e-5
    assert np.array_equal(state["be_bar"], ref["be_bar"]) and np.array_equal(state["F"], F)


def test_FeFp_plasticity_yield_stress_callable(jm, Nbatch=10):
    """tests/test_FeFp_jax.py:6-33 line for line, including the ``yield_stress`` *function* it hands to
    ``FeFpJ2Plasticity`` (numpy standing in for jax.numpy): the host side recognises the Voce law behind it."""
    E = 70e3
    nu = 0.3
    sig0 = 500.0

    b = 1000
    sigu = 750.0

    def yield_stress(p):
        return sig0 + (sigu - sig0) * (1 - np.exp(-b * p))

    elastic_model = jm.LinearElasticIsotropic(E=E, nu=nu)

    behavior = jm.FeFpJ2Plasticity(elasticity=elastic_model, yield_stress=yield_stress)
    material = jm.CUDAMaterial(behavior)
    material.set_data_manager(Nbatch)
    assert material.material_properties == {"E": E, "nu": nu, "sig0": 500.0, "sigu": 750.0, "b": 1000.0, "H": 0.0}

    eps = 2e-2

    Nsteps = 20
    dt = 0
    st = fefp.virgin_state(Nbatch)
    for t in np.linspace(0, 1.0, Nsteps)[1:]:
        F = np.zeros((Nbatch, 9))
        F[:, 0] = 1 + eps * t
        F[:, [1, 2]] = 1 - eps / 2 * t
        P, isv, Ct = material.integrate(F, dt)
        ref = fefp.integrate(F, st, PROPS)
        assert np.array_equal(P, ref["PK1"]) and np.array_equal(Ct, ref["Ct"]) and np.array_equal(isv[:, 0], ref["p"])

        material.data_manager.update()
        st = fefp.advance(ref)
    assert abs(isv[0, 0] - 1.076097e-2) < 5e-9 and abs(P[0, 0] - 473.1527) < 5e-5


@pytest.mark.parametrize("n", [1, 33, 1000, 50021])
def test_random_history_bit_exact(jm, n):
    material = make(jm, n)
    st = fefp.virgin_state(n)
    K = 4
    for k in range(1, K + 1):
        ref = compare(material, synth.defgrad(n, 0, 3e-2, k, K), st)
        material.data_manager.update()
        st = fefp.advance(ref)
    if n >= 1000:
        assert 0.2 < ref["flag"].mean() < 0.95
        assert material.last_stats.n_plastic == int(ref["flag"].sum())
        assert material.last_stats.max_iter == int(ref["n_iter"].max())


def test_initial_state_is_identity(jm):
    material = make(jm, 5, diag=False)
    s0 = material.get_initial_state_dict()
    assert np.array_equal(s0["F"], np.tile([1, 1, 1, 0, 0, 0, 0, 0, 0.0], (5, 1)))
    assert np.array_equal(s0["be_bar"], np.tile([1, 1, 1, 0, 0, 0.0], (5, 1)))
    assert np.count_nonzero(s0["PK1"]) == 0 and np.count_nonzero(s0["p"]) == 0
    # explicit (re)initialisation as the demo does (finite_strain_elastoplasticity.py:181)
    material.set_initial_state_dict({"be_bar": np.tile([1, 1, 1, 0, 0, 0.0], (5, 1))})


def test_inverted_element_flagged(jm):
    material = make(jm, 3)
    F = synth.defgrad(3, 0, 1e-2, 1, 1)
    F[1, 0] = -1.0
    with pytest.warns(jm.PerformanceWarning):
        material.integrate(F)
    assert material.diagnostics()[3].tolist() == [0, 1, 0]


def test_resident_synth_matches_oracle(jm):
    n = 4000
    material = make(jm, n)
    material.synth_gradients(seed=0, amp=3e-2, k=2, K=4)
    material.integrate_resident()
    F = synth.defgrad(n, 0, 3e-2, 2, 4)
    assert np.array_equal(material.device_view("F").cpu().numpy().T, F)
    ref = fefp.integrate(F, fefp.virgin_state(n), PROPS)
    assert np.array_equal(material.device_view("PK1").cpu().numpy().T, ref["PK1"])
    assert np.array_equal(material.device_view("Ct").cpu().numpy().T.reshape(n, 9, 9), ref["Ct"])


def test_per_point_properties(jm):
    """Heterogeneous FeFp batch: per-Gauss-point E, nu, sig0, sigu, b, H (quadrature_map.py:160-172) incl. an
    elastic class (sig0 = inf) -- exercises the PERPOINT instantiation of the kernel."""
    n = 20011
    cls = np.arange(n) % 3
    props = {
        "E": np.where(cls == 1, 90e3, 70e3), "nu": np.where(cls == 1, 0.25, 0.3),
        "sig0": np.where(cls == 2, np.inf, 500.0), "sigu": np.where(cls == 2, np.inf, np.where(cls == 1, 600.0, 750.0)),
        "b": np.where(cls == 1, 50.0, 1000.0), "H": np.where(cls == 0, 100.0, 0.0),
    }
    material = make(jm, n)
    for k, v in props.items():
        material.update_material_property(k, v)
    st = fefp.virgin_state(n)
    for k in range(1, 4):
        F = synth.defgrad(n, 7, 4e-2, k, 3)
        P, isv, Ct = material.integrate(F)
        ref = fefp.integrate(F, st, props)
        flag, n_iter, _, fail = material.diagnostics()
        assert np.array_equal(flag, ref["flag"]) and np.array_equal(n_iter, ref["n_iter"]) and fail.sum() == 0
        assert np.array_equal(P, ref["PK1"]) and np.array_equal(Ct, ref["Ct"])
        assert np.array_equal(isv[:, 1:], ref["be_bar"])
        material.data_manager.update()
        st = fefp.advance(ref)
    assert ref["flag"][cls == 2].sum() == 0 and ref["flag"][cls == 0].mean() > 0.3
