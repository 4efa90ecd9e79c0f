"""GPU parity of the Hosford behaviour (``DXM_HOSFORD_LINEAR``; matrix phase of the reference's multi-material demo,
``demos/multimaterials/multimaterials.py:245-254`` / ``IsotropicPlasticHosfordFlowLinear.mfront``) against the CPU
oracle through the public API: identical active sets, local iteration counts and failure flags; stress, state,
tangent and residual bit for bit (north-star tolerance: rtol 1e-10)."""

import numpy as np
import pytest

from oracle import hosford as ho
from oracle import small_strain as ss
from oracle import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-10
DEMO = dict(E=70e3, nu=0.3, sig0=200.0, H=10.0, a=10)  # multimaterials.py:245-254


def make(jm, props, n, diag=True):
    hard = (jm.VoceHardening(sig0=props["sig0"], sigu=props["sigu"], b=props["b"], H=props.get("H", 0.0)) if "sigu" in props
            else jm.LinearHardening(sig0=props["sig0"], H=props["H"]))
    beh = jm.GeneralIsotropicHardening(
        elasticity=jm.LinearElasticIsotropic(E=props["E"], nu=props["nu"]),
        yield_stress=hard,
        equivalent_stress=jm.Hosford(a=props["a"]),
    )
    m = jm.CUDAMaterial(beh)
    m.set_data_manager(n)
    if diag:
        m.enable_diagnostics()
    return m


def check(m, out, ref):
    flux, isv, Ct = out
    flag, n_iter, resid, fail = m.diagnostics()
    assert np.array_equal(flag, ref["flag"]) and np.array_equal(n_iter, ref["n_iter"]) and np.array_equal(fail, ref["fail"])
    np.testing.assert_allclose(flux, ref["stress"], rtol=RTOL, atol=0)
    assert np.array_equal(flux, ref["stress"]) and np.array_equal(Ct, ref["Ct"])
    assert np.array_equal(isv[:, 0], ref["p"]) and np.array_equal(isv[:, 1:], ref["epsp"])
    assert np.array_equal(resid, ref["resid"])
    s = m.last_stats
    assert s.n_plastic == int(ref["flag"].sum()) and s.n_fail == int(ref["fail"].sum())
    assert s.max_iter == int(ref["n_iter"].max()) and s.max_residual == ref["resid"].max()


@pytest.mark.parametrize("split", ["0", "1"])
@pytest.mark.parametrize("n", [129, 50_003])
def test_fused_and_split_launches_agree_with_the_oracle(jm, monkeypatch, split, n):
    """DXM_HOS_SPLIT: fused kernel (0) vs tiled kernel (1: stream a tile, CTA-local candidate queue, packed local
    solves) -- same bits either way."""
    monkeypatch.setenv("DXM_HOS_SPLIT", split)
    m = make(jm, DEMO, n)
    st = ss.zero_state(n)
    for k, amp in ((1, 2e-3), (2, 6e-3), (3, 1.25e-2)):  # few, some, most points plastic
        eps = synth.strain(n, 4, amp, 1, 1)
        out = m.integrate(eps)
        ref = ho.integrate(eps, st, DEMO)
        check(m, out, ref)
        m.data_manager.update()
        st = ss.advance(ref)
    assert m.last_stats.n_plastic > 0.5 * n


@pytest.mark.parametrize("a", [2, 6, 10, 20])
@pytest.mark.parametrize("n", [1, 129, 50_003])
def test_history_bit_exact(jm, a, n):
    props = dict(DEMO, a=a)
    m = make(jm, props, n)
    st = ss.zero_state(n)
    for k in range(1, 5):
        eps = synth.strain(n, a, 1.25e-2, k, 4)
        out = m.integrate(eps)
        ref = ho.integrate(eps, st, props)
        check(m, out, ref)
        m.data_manager.update()
        st = ss.advance(ref)
    if n > 1000:
        assert 0.3 < ref["flag"].mean() < 0.95 and ref["fail"].sum() == 0


def test_exponent_two_reproduces_the_j2_kernel(jm):
    """a = 2 is von Mises: the Hosford kernel (eigen-decomposition, 4-unknown Newton, spectral tangent) must agree
    with the closed-form J2 kernel to rounding."""
    n = 20_000
    props = dict(DEMO, a=2)
    mh = make(jm, props, n, diag=False)
    mj = jm.CUDAMaterial(jm.vonMisesIsotropicHardening(
        elasticity=jm.LinearElasticIsotropic(E=props["E"], nu=props["nu"]),
        yield_stress=jm.LinearHardening(sig0=props["sig0"], H=props["H"])))
    mj.set_data_manager(n)
    for k in range(1, 4):
        eps = synth.strain(n, 1, 1.25e-2, k, 3)
        fh, ih, ch = [x.copy() for x in mh.integrate(eps)]
        fj, ij, cj = mj.integrate(eps)
        np.testing.assert_allclose(fh, fj, rtol=1e-12, atol=1e-10)
        np.testing.assert_allclose(ih, ij, rtol=1e-10, atol=1e-16)
        np.testing.assert_allclose(ch, cj, rtol=1e-9, atol=1e-8 * props["E"])
        assert mh.last_stats.n_plastic == mj.last_stats.n_plastic
        mh.data_manager.update()
        mj.data_manager.update()


def test_per_point_properties_degenerate_spectra_and_large_steps(jm):
    n = 30_001
    rng = np.random.default_rng(5)
    E = rng.uniform(60e3, 90e3, n)
    s0 = np.where(np.arange(n) % 3 == 0, 200.0, 260.0)
    props = dict(E=E, nu=0.3, sig0=s0, H=10.0, a=10)
    m = make(jm, dict(DEMO), n)
    m.update_material_property("E", E)
    m.update_material_property("sig0", s0)
    eps = synth.strain(n, 9, 0.1, 1, 1)  # up to ~35 x the yield strain
    eps[:100, 1] = eps[:100, 2]  # repeated principal strains
    eps[:100, 3:] = 0.0
    eps[100:200] = 0.0  # nothing happens
    eps[200:300, :3] = eps[200:300, :1]  # hydrostatic
    eps[200:300, 3:] = 0.0
    out = m.integrate(eps)
    ref = ho.integrate(eps, ss.zero_state(n), props)
    check(m, out, ref)
    assert ref["fail"].sum() == 0 and ref["flag"][100:300].sum() == 0 and ref["flag"][:100].sum() > 50


def test_resident_path_device_tangent_and_exponent_update(jm):
    n = 10_000
    m = make(jm, DEMO, n, diag=False)
    eps = synth.strain(n, 2, 1.25e-2, 1, 1)
    flux, isv, Ct = [x.copy() for x in m.integrate(eps)]
    import torch

    m.gradient_buffer().copy_(torch.as_tensor(eps.T.copy(), device="cuda"))
    m.integrate_resident()
    assert np.array_equal(m.device_view("stress").T.cpu().numpy(), flux)
    assert np.array_equal(m.device_tangent().T.cpu().numpy().reshape(n, 6, 6), Ct)
    m.update_material_property("a", 6)
    f6, _, _ = m.integrate(eps)
    ref = ho.integrate(eps, ss.zero_state(n), dict(DEMO, a=6))
    assert np.array_equal(f6, ref["stress"])
    from dolfinx_materials_b200._lib import DxmError

    with pytest.raises(DxmError):
        m.update_material_property("a", 7)
    with pytest.raises(DxmError):
        m.update_material_property("a", np.full(n, 10.0))


def test_nan_gradient_is_reported_as_a_failed_point(jm):
    n = 64
    m = make(jm, DEMO, n)
    eps = synth.strain(n, 0, 1e-2, 1, 1)
    eps[7, 0] = np.nan
    import warnings

    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        out = m.integrate(eps)
    ref = ho.integrate(eps, ss.zero_state(n), DEMO)
    assert m.last_stats.n_fail == 1 == int(ref["fail"].sum())
    assert any("failed" in str(x.message) for x in w)
    ok = np.arange(n) != 7
    assert np.array_equal(out[0][ok], ref["stress"][ok]) and np.array_equal(out[2][ok], ref["Ct"][ok])


@pytest.mark.parametrize("split", ["0", "1"])
@pytest.mark.parametrize("a", [2, 10])
def test_voce_hardening_behind_the_hosford_criterion(jm, monkeypatch, split, a):
    """jaxmat's GeneralIsotropicHardening(elastic, Voce yield stress, ...) with the Hosford norm: bit-exact vs the
    oracle; per-point saturation stress; a = 2 agrees with the J2 + Voce kernel to rounding."""
    monkeypatch.setenv("DXM_HOS_SPLIT", split)
    props = dict(E=70e3, nu=0.3, sig0=350.0, sigu=500.0, b=1e3, H=25.0, a=a)
    n = 40_001
    m = make(jm, props, n)
    sigu = np.where(np.arange(n) % 2 == 0, 500.0, 430.0)
    m.update_material_property("sigu", sigu)
    st = ss.zero_state(n)
    for k in range(1, 4):
        eps = synth.strain(n, 6, 1.25e-2, k, 3)
        out = m.integrate(eps)
        ref = ho.integrate(eps, st, dict(props, sigu=sigu))
        check(m, out, ref)
        m.data_manager.update()
        st = ss.advance(ref)
    assert 0.3 < ref["flag"].mean() < 0.95 and ref["n_iter"].max() <= 8
    if a == 2:
        j = jm.CUDAMaterial(jm.vonMisesIsotropicHardening(
            elasticity=jm.LinearElasticIsotropic(E=props["E"], nu=props["nu"]),
            yield_stress=jm.VoceHardening(sig0=350.0, sigu=500.0, b=1e3, H=25.0)))
        j.set_data_manager(n)
        j.update_material_property("sigu", sigu)
        h2 = make(jm, props, n, diag=False)
        h2.update_material_property("sigu", sigu)
        eps = synth.strain(n, 6, 1.25e-2, 1, 1)
        fj, _, cj = j.integrate(eps)
        fh, _, ch = h2.integrate(eps)
        np.testing.assert_allclose(fh, fj, rtol=1e-11, atol=1e-9)
        np.testing.assert_allclose(ch, cj, rtol=1e-8, atol=1e-7 * props["E"])
