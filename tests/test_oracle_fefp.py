"""Self-consistency of the finite-strain FeFp oracle (CPU only).  Parity with jaxmat itself is
unpinned (un-vendored dependency; the reference test tests/test_FeFp_jax.py asserts nothing) -- these
tests check the restatement against (i) the reference test's own load path, whose end values were
obtained independently by the survey prototype (p = 1.076097e-2, P11 = 473.1527,
elastic for the first 5 steps), (ii) an independent solve of the reference's 7-unknown formulation,
(iii) finite-difference tangents, yield consistency and det(be_bar) = 1."""
import numpy as np
import pytest

from oracle import fefp, synth

PROPS = dict(E=70e3, nu=0.3, sig0=500.0, sigu=750.0, b=1000.0)  # tests/test_FeFp_jax.py:7-15
MU, KAPPA = 70e3 / 2 / 1.3, 70e3 / (3 * 0.4)


def mat9(v):
    return np.array([[v[fefp.IDX9[i][j]] for j in range(3)] for i in range(3)])


def mat6(b):
    r = 2 ** -0.5
    return np.array([[b[0], b[3] * r, b[4] * r], [b[3] * r, b[1], b[5] * r], [b[4] * r, b[5] * r, b[2]]])


def test_rcbrt_c():
    x = np.concatenate([np.linspace(0.2, 5, 100001), np.logspace(-300, 300, 5001), [1.0, 8.0, 27.0, 1e-3]])
    y = fefp.rcbrt_c(x)
    ref = 1.0 / np.cbrt(x)
    assert (np.abs(y - ref) <= 3.0 * np.spacing(ref)).all()
    assert fefp.rcbrt_c(1.0) == 1.0 and np.isnan(fefp.rcbrt_c(-1.0)) and np.isnan(fefp.rcbrt_c(0.0))


def test_reference_test_path():
    """tests/test_FeFp_jax.py:20-33: F = diag(1 + eps t, 1 - eps t/2, 1 - eps t/2), eps = 2e-2, 19 steps."""
    n, eps = 10, 2e-2
    st = fefp.virgin_state(n)
    flags = []
    for t in np.linspace(0, 1.0, 20)[1:]:
        F = np.zeros((n, 9))
        F[:, 0] = 1 + eps * t
        F[:, [1, 2]] = 1 - eps / 2 * t
        out = fefp.integrate(F, st, PROPS)
        st = fefp.advance(out)
        flags.append(int(out["flag"][0]))
        assert out["fail"].sum() == 0
    assert flags == [0] * 5 + [1] * 14
    assert abs(out["p"][0] - 1.076097e-2) < 5e-9
    assert abs(out["PK1"][0, 0] - 473.1527) < 5e-5
    assert np.ptp(out["PK1"], axis=0).max() == 0  # identical points give identical results
    assert abs(np.linalg.det(mat6(out["be_bar"][0])) - 1) < 1e-12


@pytest.fixture(scope="module")
def history():
    n, K = 400, 4
    st = fefp.virgin_state(n)
    for k in range(1, K):
        st = fefp.advance(fefp.integrate(synth.defgrad(n, 0, 6e-2, k, K), st, PROPS))
    F = synth.defgrad(n, 0, 6e-2, K, K)
    return st, F, fefp.integrate(F, st, PROPS)


def test_tangent_matches_finite_differences(history):
    st, F, out = history
    assert 0.5 < out["flag"].mean() < 0.99 and out["fail"].sum() == 0
    Ct, h = out["Ct"], 1e-7
    fd = np.zeros_like(Ct)
    for b in range(9):
        Fp, Fm = F.copy(), F.copy()
        Fp[:, b] += h
        Fm[:, b] -= h
        fd[:, :, b] = (fefp.integrate(Fp, st, PROPS)["PK1"] - fefp.integrate(Fm, st, PROPS)["PK1"]) / (2 * h)
    err = np.abs(fd - Ct).max(axis=(1, 2)) / np.abs(Ct).max(axis=(1, 2))
    assert err.max() < 2e-8


def test_yield_consistency_det_and_symmetry(history):
    st, F, out = history
    for i in range(F.shape[0]):
        tau = mat9(out["PK1"][i]) @ mat9(F[i]).T
        assert np.abs(tau - tau.T).max() < 1e-9  # Kirchhoff stress is symmetric
        if out["flag"][i]:
            s = tau - np.trace(tau) / 3 * np.eye(3)
            vm = np.sqrt(1.5 * (s * s).sum())
            sy = 500 + 250 * (1 - np.exp(-1000 * out["p"][i]))
            assert abs(vm - sy) < 1e-8 * sy
            assert out["p"][i] > st["p"][i]
        assert abs(np.linalg.det(mat6(out["be_bar"][i])) - 1) < 2e-12


def test_reduced_solve_equals_seven_unknown_formulation(history):
    """unknowns (dp, be_bar), residuals = yield condition and
    dev(be - be_tr) + 2/3 dp tr(be) n + 1 (det be - 1) = 0, solved by scipy from the trial state."""
    from scipy.optimize import fsolve

    st, F, out = history
    worst = 0.0
    for i in np.flatnonzero(out["flag"])[:25]:
        f = mat9(F[i]) @ np.linalg.inv(mat9(st["F"][i]))
        fb = f * np.linalg.det(f) ** (-1 / 3)
        Btr = fb @ mat6(st["be_bar"][i]) @ fb.T
        p_old = st["p"][i]

        def res(x):
            be = np.array([[x[1], x[4], x[5]], [x[4], x[2], x[6]], [x[5], x[6], x[3]]])
            s = MU * (be - np.trace(be) / 3 * np.eye(3))
            seq = np.sqrt(1.5 * (s * s).sum())
            fy = seq - (500 + 250 * (1 - np.exp(-1000 * (p_old + x[0]))))
            R = (be - Btr) - np.trace(be - Btr) / 3 * np.eye(3) + 2 / 3 * x[0] * np.trace(be) * 1.5 * s / seq
            R = R + np.eye(3) * (np.linalg.det(be) - 1)
            return np.array([fy / 70e3, R[0, 0], R[1, 1], R[2, 2], R[0, 1], R[0, 2], R[1, 2]])

        x0 = np.array([1e-4, Btr[0, 0], Btr[1, 1], Btr[2, 2], Btr[0, 1], Btr[0, 2], Btr[1, 2]])
        x = fsolve(res, x0, xtol=1e-14)
        assert np.abs(res(x)).max() < 1e-12
        be = mat6(out["be_bar"][i])
        mine = np.array([out["p"][i] - p_old, be[0, 0], be[1, 1], be[2, 2], be[0, 1], be[0, 2], be[1, 2]])
        worst = max(worst, np.abs(mine - x).max())
    assert worst < 1e-12


def test_elastic_step_keeps_trial_and_is_hyperelastic():
    n = 50
    F = synth.defgrad(n, 3, 5e-3, 1, 1)
    out = fefp.integrate(F, fefp.virgin_state(n), PROPS)
    assert out["flag"].sum() == 0 and out["n_iter"].max() == 0
    for i in range(n):
        Fm = mat9(F[i])
        J = np.linalg.det(Fm)
        b = Fm @ Fm.T * J ** (-2 / 3)
        tau = MU * (b - np.trace(b) / 3 * np.eye(3)) + KAPPA / 2 * (J * J - 1) * np.eye(3)
        P = tau @ np.linalg.inv(Fm).T
        np.testing.assert_allclose(mat9(out["PK1"][i]), P, rtol=1e-10, atol=1e-9)
        np.testing.assert_allclose(mat6(out["be_bar"][i]), b, rtol=1e-13, atol=1e-15)


def test_inverted_element_is_flagged():
    F = synth.defgrad(3, 0, 1e-2, 1, 1)
    F[1, 0] = -1.0  # det F < 0 relative to the identity reference
    out = fefp.integrate(F, fefp.virgin_state(3), PROPS)
    assert out["fail"].tolist() == [0, 1, 0]


def test_singular_elastic_state_is_a_failed_point():
    """be_bar = 0 (a state Function that was never initialised to the identity) gives PK1 = 0 with every finiteness
    check green; det(be_bar_old) <= 0 is therefore counted as a failure, in the oracle and the kernel alike."""
    n = 4
    F = np.tile([1.0, 1, 1, 0.01, 0, 0, 0, 0, 0], (n, 1))
    st = fefp.virgin_state(n)
    st["be_bar"][1] = 0.0  # singular
    st["be_bar"][2] = [-1.0, 1, 1, 0, 0, 0]  # inverted
    out = fefp.integrate(F, st, PROPS)
    assert out["fail"].tolist() == [0, 1, 1, 0]
    assert np.all(out["PK1"][1] == 0.0) and np.abs(out["PK1"][0]).max() > 100.0  # the silent result this guards against
    from oracle import cport

    assert cport.fefp(F, st, PROPS)["fail"].tolist() == [0, 1, 1, 0]


def test_history_matches_reference_protocol_run():
    """tests/golden/fefp_history.npz: the reference's own Material.integrate / _vmap / DataManager drove a
    per-point FeFp material over a 3-increment history (make_golden.py); the batched oracle with explicit
    state carry must reproduce it: bit for bit with every fma split (the fixture holds the round-1 arithmetic), to
    rtol 1e-10 with identical active sets / iteration counts in the fused canonical arithmetic.  Pins protocol +
    regression, not jaxmat parity."""
    import os

    from golden_check import close, same_active_set
    from oracle import canon

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fefp_history.npz"))
    props = dict(zip([str(k) for k in g["props_keys"]], [float(v) for v in g["props_vals"]]))
    n = g["F1"].shape[0]
    st = st_u = fefp.virgin_state(n)
    k = 1
    while f"F{k}" in g:
        with canon.unfused():
            ref = fefp.integrate(g[f"F{k}"], st_u, props)
        assert np.array_equal(ref["PK1"], g[f"flux{k}"])
        assert np.array_equal(ref["p"], g[f"isv{k}"][:, 0]) and np.array_equal(ref["be_bar"], g[f"isv{k}"][:, 1:])
        assert np.array_equal(ref["Ct"], g[f"Ct{k}"])
        out = fefp.integrate(g[f"F{k}"], st, props)
        same_active_set(out, ref)
        close(out["PK1"], g[f"flux{k}"], "PK1")
        close(out["p"], g[f"isv{k}"][:, 0], "p")
        close(out["be_bar"], g[f"isv{k}"][:, 1:], "be_bar")
        close(out["Ct"], g[f"Ct{k}"], "Ct")
        st, st_u = fefp.advance(out), fefp.advance(ref)
        k += 1
    assert out["flag"].any()
