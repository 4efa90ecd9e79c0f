"""Pins the small-strain oracle (CPU only): against the golden vectors produced by the reference's
own code (tests/golden/make_golden.py), the closed form of the in-tree MFront source, the analytic
limit the reference's test asserts, and self-consistency (finite-difference tangent, yield surface)."""
import os

import numpy as np
import pytest

from oracle import canon, synth
from oracle import small_strain as ss

GOLD = os.path.join(os.path.dirname(__file__), "golden")
VOCE = dict(E=70e3, nu=0.3, sig0=350.0, sigu=500.0, b=1e3)


def test_elastic_matches_reference_run():
    """LinearElasticIsotropic through the reference's Material.integrate (golden, reference-executed)."""
    g = np.load(os.path.join(GOLD, "elastic_reference.npz"))
    n = g["eps1"].shape[0]
    st = ss.zero_state(n)
    for k in (1, 2):
        out = ss.integrate(g[f"eps{k}"], st, ss.elastic_props(70e3, 0.3))
        # the reference computes sigma = C @ eps (total form); the oracle accumulates increments
        np.testing.assert_allclose(out["stress"], g[f"flux{k}"], rtol=1e-13, atol=1e-10)
        assert np.array_equal(out["Ct"], g[f"Ct{k}"])  # exactly C
        assert not out["flag"].any()
        st = ss.advance(out)
        np.testing.assert_allclose(st["stress"], g[f"s0_stress_after_update{k}"], rtol=1e-13, atol=1e-10)
    out = ss.integrate(np.array([[1e-3, 0, 0, 0, 0, 0.0]]), ss.zero_state(1), ss.elastic_props(70e3, 0.0))
    assert np.allclose(out["stress"][0, :3], 70e3 * np.array([1e-3, 0, 0]))  # test_initialization.py:131-153
    assert np.allclose(out["stress"], g["flux_nu0"])


@pytest.mark.parametrize("name", ["j2_voce_history.npz", "j2_linear_history.npz"])
def test_history_matches_reference_protocol_run(name):
    """Batched oracle + state carry == the reference's per-point _vmap/DataManager protocol.  The fixture holds the
    round-1 (un-fused) arithmetic: reproduced bit for bit with every fma split, and to rtol 1e-10 with identical
    active sets / iteration counts by the fused canonical arithmetic (tests/golden_check.py)."""
    from golden_check import close, same_active_set
    from oracle import canon

    g = np.load(os.path.join(GOLD, name))
    props = dict(zip([str(k) for k in g["props_keys"]], [float(v) for v in g["props_vals"]]))
    n = g["eps1"].shape[0]
    st = st_u = ss.zero_state(n)
    k = 1
    while f"eps{k}" in g:
        with canon.unfused():
            ref = ss.integrate(g[f"eps{k}"], st_u, props)
        assert np.array_equal(ref["stress"], g[f"flux{k}"])
        assert np.array_equal(ref["p"], g[f"isv{k}"][:, 0])
        assert np.array_equal(ref["epsp"], g[f"isv{k}"][:, 1:])
        assert np.array_equal(ref["Ct"], g[f"Ct{k}"])
        out = ss.integrate(g[f"eps{k}"], st, props)
        same_active_set(out, ref)
        close(out["stress"], g[f"flux{k}"], "stress")
        close(out["p"], g[f"isv{k}"][:, 0], "p")
        close(out["epsp"], g[f"isv{k}"][:, 1:], "epsp")
        close(out["Ct"], g[f"Ct{k}"], "Ct")
        st, st_u = ss.advance(out), ss.advance(ref)
        k += 1
    assert out["flag"].any() and not out["flag"].all()


def _mfront_closed_form(eel_old, p_old, deto, young, nu, H, s0):
    """Literal transcription of tests/mfront/IsotropicLinearHardeningPlasticity.mfront:49-77 for one
    point, with MFront's tensor objects written out in Mandel notation."""
    lam = young * nu / ((1 + nu) * (1 - 2 * nu))
    mu = young / (2 * (1 + nu))
    Id = np.array([1, 1, 1, 0, 0, 0.0])
    IxI = np.outer(Id, Id)
    I4 = np.eye(6)
    M = 1.5 * (I4 - IxI / 3)
    eel = eel_old + deto
    se = 2 * mu * (eel - eel[:3].sum() / 3 * Id)
    seq = np.sqrt(1.5 * se @ se)
    if seq - s0 - H * p_old > 0:
        n = 3 * se / (2 * seq)
        cste = 1 / (H + 3 * mu)
        dp = (seq - s0 - H * p_old) * cste
        eel = eel - dp * n
        Dt = lam * IxI + 2 * mu * I4 - 4 * mu * mu * (dp / seq * (M - np.outer(n, n)) + cste * np.outer(n, n))
    else:
        dp = 0.0
        Dt = lam * IxI + 2 * mu * I4
    sig = lam * eel[:3].sum() * Id + 2 * mu * eel
    return sig, eel, p_old + dp, Dt


def test_linear_hardening_equals_mfront_closed_form():
    props = dict(E=70e3, nu=0.3, sig0=250.0, H=5e3)
    n, K = 200, 3
    st = ss.zero_state(n)
    eel = np.zeros((n, 6))
    p = np.zeros(n)
    eps_old = np.zeros((n, 6))
    for k in range(1, K + 1):
        eps = synth.strain(n, 2, 1.5e-2, k, K)
        out = ss.integrate(eps, st, props)
        for i in range(n):
            sig, eel[i], p[i], Dt = _mfront_closed_form(eel[i], p[i], eps[i] - eps_old[i], 70e3, 0.3, 5e3, 250.0)
            np.testing.assert_allclose(out["stress"][i], sig, rtol=1e-11, atol=1e-9)
            np.testing.assert_allclose(out["Ct"][i], Dt, rtol=1e-11, atol=1e-7)
            assert abs(out["p"][i] - p[i]) <= 1e-15 + 1e-11 * p[i]
        eps_old = eps
        st = ss.advance(out)
    assert out["flag"].sum() > 20


def test_plane_strain_uniaxial_limit():
    """tests/mfront/test_elastoplasticity.py:16-36: E=70e3, nu=0.3, H=1e-6, sig0=250, 50 steps to 2 %:
    Stress[:3] -> 2/sqrt(3) [sig0, 0, sig0/2] at rtol 1e-2.  The FE solve reduces, for the uniform
    1-element problem, to finding eyy with sigma_yy = 0 at every step."""
    from scipy.optimize import brentq

    props = dict(E=70e3, nu=0.3, sig0=250.0, H=1e-6)
    st = ss.zero_state(1)
    for exx in np.linspace(0, 2e-2, 51)[1:]:
        def syy(eyy):
            return ss.integrate(np.array([[exx, eyy, 0, 0, 0, 0.0]]), st, props)["stress"][0, 1]

        eyy = brentq(syy, -exx, exx, xtol=1e-16, rtol=1e-15)
        out = ss.integrate(np.array([[exx, eyy, 0, 0, 0, 0.0]]), st, props)
        st = ss.advance(out)
    assert np.allclose(out["stress"][0, :3], 2 / np.sqrt(3) * np.array([250.0, 0, 125.0]), rtol=1e-2, atol=1e-8)


def _history(props, n, amp, K, seed=0):
    st = ss.zero_state(n)
    for k in range(1, K):
        st = ss.advance(ss.integrate(synth.strain(n, seed, amp, k, K), st, props))
    return st, synth.strain(n, seed, amp, K, K)


def test_voce_tangent_yield_and_iterations():
    n = 3000
    st, eps = _history(VOCE, n, 1.25e-2, 4)
    out = ss.integrate(eps, st, VOCE)
    pl = out["flag"] == 1
    assert 0.4 < pl.mean() < 0.8 and out["fail"].sum() == 0
    assert out["n_iter"][~pl].max() == 0 and 1 <= out["n_iter"][pl].min() and out["n_iter"].max() <= 6
    # yield consistency f(sigma, p) = 0 on the active set; dp >= 0
    sig = out["stress"]
    s = sig.copy()
    s[:, :3] -= sig[:, :3].mean(1, keepdims=True)
    seq = np.sqrt(1.5 * (s * s).sum(1))
    sy = 350.0 + 150.0 * (1 - np.exp(-1e3 * out["p"]))
    assert np.abs(seq - sy)[pl].max() < 1e-9 * 350
    assert (out["p"] - st["p"]).min() >= 0
    # consistent tangent vs central differences; symmetric; elastic points return exactly C
    Ct = out["Ct"]
    assert np.array_equal(Ct, Ct.transpose(0, 2, 1))
    h = 1e-7
    fd = np.zeros_like(Ct)
    for i in range(6):
        ep, em = eps.copy(), eps.copy()
        ep[:, i] += h
        em[:, i] -= h
        fd[:, :, i] = (ss.integrate(ep, st, VOCE)["stress"] - ss.integrate(em, st, VOCE)["stress"]) / (2 * h)
    err = np.abs(fd - Ct).max(axis=(1, 2)) / np.abs(Ct).max(axis=(1, 2))
    assert err.max() < 1e-6
    lam, mu = canon.lame(70e3, 0.3)
    C = 2 * mu * np.eye(6)
    C[:3, :3] += lam
    assert np.array_equal(Ct[~pl], np.broadcast_to(C, Ct[~pl].shape))


def test_newton_matches_bracketing_solver():
    """The local Newton solution equals an independent bracketing solve of the same scalar equation."""
    from scipy.optimize import brentq

    n = 50
    st, eps = _history(VOCE, n, 2e-2, 2, seed=4)
    out = ss.integrate(eps, st, VOCE)
    lam, mu = canon.lame(70e3, 0.3)
    for i in np.flatnonzero(out["flag"]):
        de = eps[i] - st["strain"][i]
        sig_tr = st["stress"][i] + lam * de[:3].sum() * np.array([1, 1, 1, 0, 0, 0.0]) + 2 * mu * de
        s = sig_tr.copy()
        s[:3] -= sig_tr[:3].mean()
        seq = np.sqrt(1.5 * s @ s)
        r = lambda dp: seq - 3 * mu * dp - (350.0 + 150.0 * (1 - np.exp(-1e3 * (st["p"][i] + dp))))  # noqa: E731
        dp = brentq(r, 0.0, seq / (3 * mu), xtol=1e-18, rtol=1e-15)
        assert abs((out["p"][i] - st["p"][i]) - dp) <= 1e-10 * dp  # north-star rtol


def test_cap_and_nonfinite_set_fail():
    out = ss.integrate(np.array([[np.nan, 0, 0, 0, 0, 0.0], [1e-2, 0, 0, 0, 0, 0]]), ss.zero_state(2), VOCE)
    assert out["fail"].tolist() == [1, 0]
    out = ss.integrate(np.array([[1e-2, 0, 0, 0, 0, 0.0]]), ss.zero_state(1), VOCE, newton_cap=1)
    assert out["fail"][0] == 1 and out["n_iter"][0] == 1


def test_mixed_per_point_properties_equal_separate_batches():
    """Heterogeneous batch (cfg4) == the same points integrated class by class; disjoint supports as in
    tests/mfront/test_multimaterials.py:163-172."""
    n = 600
    cls = np.arange(n) % 3
    props = {
        "E": np.where(cls == 1, 90e3, 70e3), "nu": np.where(cls == 1, 0.25, 0.3),
        "sig0": np.where(cls == 2, np.inf, 200.0), "H": np.where(cls == 0, 10.0, 0.0),
        "sigu": np.where(cls == 1, 300.0, np.where(cls == 2, np.inf, 200.0)), "b": np.where(cls == 1, 10.0, 0.0),
    }
    eps = synth.strain(n, 9, 1.25e-2, 1, 1)
    out = ss.integrate(eps, ss.zero_state(n), props)
    uni = [dict(E=70e3, nu=0.3, sig0=200.0, H=10.0), dict(E=90e3, nu=0.25, sig0=200.0, sigu=300.0, b=10.0),
           ss.elastic_props(70e3, 0.3)]
    for c in range(3):
        sel = cls == c
        ref = ss.integrate(eps[sel], ss.zero_state(int(sel.sum())), uni[c])
        for key in ("stress", "p", "epsp", "Ct", "flag", "n_iter"):
            assert np.array_equal(out[key][sel], ref[key])
    assert out["flag"][cls == 2].sum() == 0 and out["n_iter"][cls == 0].max() == 0
