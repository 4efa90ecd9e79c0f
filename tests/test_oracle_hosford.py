"""The Hosford oracle (canonical C restatement of demos/multimaterials/IsotropicPlasticHosfordFlowLinear.mfront)
checked without a GPU: a = 2 is von Mises (the pinned J2 oracle), the solution satisfies an independently written
statement of the implicit system, the tangent is the derivative of the stress, known yield points of the criterion
(uniaxial R0, pure shear R0 / (2^(a-1) + 1)^(1/a)), degenerate spectra, very large steps."""

import numpy as np
import pytest

from oracle import hosford as ho
from oracle import small_strain as ss
from oracle import synth

DEMO = dict(E=70e3, nu=0.3, sig0=200.0, H=10.0, a=10)  # multimaterials.py:245-254


def history(props, n, amp, K, seed=0):
    st = ss.zero_state(n)
    outs = []
    for k in range(1, K + 1):
        out = ho.integrate(synth.strain(n, seed, amp, k, K), st, props)
        outs.append((st, out))
        st = ss.advance(out)
    return outs


def test_exponent_two_is_von_mises():
    n = 3000
    p2 = dict(DEMO, a=2)
    st_h, st_j = ss.zero_state(n), ss.zero_state(n)
    for k in range(1, 5):
        eps = synth.strain(n, 1, 1.25e-2, k, 4)
        h = ho.integrate(eps, st_h, p2)
        j = ss.integrate(eps, st_j, {k_: v for k_, v in p2.items() if k_ != "a"})
        assert np.array_equal(h["flag"], j["flag"])
        np.testing.assert_allclose(h["stress"], j["stress"], rtol=1e-12, atol=1e-10)
        np.testing.assert_allclose(h["p"], j["p"], rtol=1e-11, atol=1e-18)
        np.testing.assert_allclose(h["epsp"], j["epsp"], rtol=1e-10, atol=1e-16)
        np.testing.assert_allclose(h["Ct"], j["Ct"], rtol=1e-9, atol=1e-8 * DEMO["E"])
        st_h, st_j = ss.advance(h), ss.advance(j)
    assert h["flag"].mean() > 0.5


@pytest.mark.parametrize("a", [4, 6, 10, 20])
def test_solution_satisfies_the_independent_implicit_system(a):
    props = dict(DEMO, a=a)
    n = 60
    lam = props["E"] * props["nu"] / (1 + props["nu"]) / (1 - 2 * props["nu"])
    mu = props["E"] / 2 / (1 + props["nu"])
    C = 2 * mu * np.eye(6)
    C[:3, :3] += lam
    seen = 0
    for st, out in history(props, n, 1.0e-2, 3, seed=a):
        assert out["fail"].sum() == 0
        sig_tr = st["stress"] + (out["strain"] - st["strain"]) @ C
        for i in range(n):
            phi = ho.sigma_eq(out["stress"][i], a)
            sy = ho.yield_stress(out["p"][i], props)
            if out["flag"][i]:
                seen += 1
                r6, r1 = ho.implicit_residual(out["stress"][i], out["p"][i] - st["p"][i], sig_tr[i], st["p"][i], props)
                assert np.abs(r6).max() < 2e-6 * props["sig0"]  # limited by the finite-difference flow direction
                assert abs(r1) < 1e-9 * props["sig0"]
                assert out["p"][i] > st["p"][i]
            else:
                assert phi <= sy * (1 + 1e-12)
                assert np.array_equal(out["stress"][i], sig_tr[i]) or np.allclose(out["stress"][i], sig_tr[i], rtol=1e-15)
                assert np.array_equal(out["Ct"][i], C)
        # additive split and plastic incompressibility
        np.testing.assert_allclose(out["stress"], (out["strain"] - out["epsp"]) @ C, rtol=1e-9, atol=1e-9 * props["sig0"])
        assert np.abs(out["epsp"][:, :3].sum(axis=1)).max() < 1e-15
    assert seen > n


@pytest.mark.parametrize("a", [6, 10])
def test_tangent_is_the_derivative_of_the_stress(a):
    props = dict(DEMO, a=a)
    n = 40
    (st, out) = history(props, n, 1.0e-2, 2, seed=7)[-1]
    eps = out["strain"]
    h = 1e-8
    J = np.zeros((n, 6, 6))
    for i in range(6):
        d = np.zeros(6)
        d[i] = h
        plus = ho.integrate(eps + d, st, props)
        minus = ho.integrate(eps - d, st, props)
        ok = (plus["flag"] == out["flag"]) & (minus["flag"] == out["flag"])
        J[:, :, i] = (plus["stress"] - minus["stress"]) / (2 * h)
        J[~ok, :, i] = np.nan
    good = ~np.isnan(J).any(axis=(1, 2))
    assert good.sum() > n // 2 and out["flag"][good].sum() > 5
    err = np.abs(out["Ct"][good] - J[good]).max(axis=(1, 2)) / props["E"]
    assert err.max() < 2e-6, err.max()
    np.testing.assert_array_equal(out["Ct"], np.swapaxes(out["Ct"], 1, 2))  # symmetric by construction


def drive(props, direction, emax, steps):
    st = ss.zero_state(1)
    for k in range(1, steps + 1):
        out = ho.integrate((emax * k / steps) * np.asarray(direction, dtype=float)[None, :], st, props)
        assert out["fail"].sum() == 0
        st = ss.advance(out)
    return out


def test_known_yield_points_of_the_criterion():
    a = 10
    props = dict(E=70e3, nu=0.3, sig0=200.0, H=1e-6, a=a)
    # pure shear (Mandel component 3 = sqrt2 * eps_12): sigma_eq = tau (2^(a-1) + 1)^(1/a)
    out = drive(props, [0, 0, 0, 1, 0, 0], 2e-2, 40)
    tau = out["stress"][0, 3] / np.sqrt(2.0)
    assert tau == pytest.approx(200.0 / (2.0 ** (a - 1) + 1.0) ** (1.0 / a), rel=1e-6)
    # isochoric uniaxial extension: stress deviator (2,-1,-1) s/3 -> sigma_eq = s, as for von Mises
    out = drive(props, [1, -0.5, -0.5, 0, 0, 0], 2e-2, 40)
    s = out["stress"][0]
    assert s[0] - s[1] == pytest.approx(200.0, rel=1e-6) and s[1] == pytest.approx(s[2], abs=1e-9)
    # the two repeated principal stresses do not break the tangent (exact divided differences)
    assert np.isfinite(out["Ct"]).all() and out["flag"][0] == 1
    ct = out["Ct"][0]
    assert ct[3, 3] == pytest.approx(ct[4, 4], rel=1e-9)  # shear planes 12 and 13 are equivalent
    # hydrostatic loading never yields
    out = drive(props, [1, 1, 1, 0, 0, 0], 5e-2, 3)
    assert out["flag"][0] == 0 and out["p"][0] == 0.0


def test_tangent_at_a_repeated_eigenvalue_matches_finite_differences():
    props = dict(DEMO, a=10)
    st = ss.zero_state(1)
    eps = np.array([[8e-3, -3e-3, -3e-3, 0, 0, 0]])
    out = ho.integrate(eps, st, props)
    assert out["flag"][0] == 1
    h = 1e-8
    J = np.zeros((6, 6))
    for i in range(6):
        d = np.zeros((1, 6))
        d[0, i] = h
        J[:, i] = (ho.integrate(eps + d, st, props)["stress"][0] - ho.integrate(eps - d, st, props)["stress"][0]) / (2 * h)
    assert np.abs(out["Ct"][0] - J).max() / props["E"] < 2e-6


@pytest.mark.parametrize("a", [2, 6, 10, 20])
def test_very_large_steps_converge(a):
    props = dict(DEMO, a=a)
    n = 4000
    out = ho.integrate(synth.strain(n, 11, 0.2, 1, 1), ss.zero_state(n), props)  # up to ~70 x the yield strain
    assert out["fail"].sum() == 0 and out["flag"].mean() > 0.9
    assert out["n_iter"].max() <= 12
    phi = np.array([ho.sigma_eq(s, a) for s in out["stress"][:200]])
    sy = props["sig0"] + props["H"] * out["p"][:200]
    np.testing.assert_allclose(phi[out["flag"][:200] > 0], sy[out["flag"][:200] > 0], rtol=1e-10)


def test_per_point_properties_match_uniform_runs():
    n = 500
    eps = synth.strain(n, 2, 1.0e-2, 1, 1)
    E = np.where(np.arange(n) % 2 == 0, 70e3, 90e3)
    s0 = np.where(np.arange(n) % 3 == 0, 200.0, 260.0)
    mixed = ho.integrate(eps, ss.zero_state(n), dict(E=E, nu=0.3, sig0=s0, H=10.0, a=10))
    for e in (70e3, 90e3):
        for s in (200.0, 260.0):
            sel = (E == e) & (s0 == s)
            uni = ho.integrate(eps[sel], ss.zero_state(int(sel.sum())), dict(E=e, nu=0.3, sig0=s, H=10.0, a=10))
            assert np.array_equal(uni["stress"], mixed["stress"][sel]) and np.array_equal(uni["Ct"], mixed["Ct"][sel])


def _mandel_deviators(S):
    r2 = np.sqrt(2.0)
    S = S - np.eye(3) * (np.trace(S, axis1=1, axis2=2) / 3.0)[:, None, None]
    return S, np.stack([S[:, 0, 0], S[:, 1, 1], S[:, 2, 2], r2 * S[:, 0, 1], r2 * S[:, 0, 2], r2 * S[:, 1, 2]], 1)


def test_noniterative_eigen_decomposition_is_backward_stable():
    """The update's eigen-decomposition (isolated root of the characteristic cubic -> cross product -> one Jacobi
    rotation in the normal plane) over random, nearly degenerate, pure-shear-like and exactly diagonal / axisymmetric
    spectra: residual, orthogonality and eigenvalues (vs LAPACK) to a few ulp of |A|."""
    from oracle import cport

    rng = np.random.default_rng(1)
    n = 50000
    sym = lambda M: (M + np.swapaxes(M, 1, 2)) / 2
    Q, _ = np.linalg.qr(rng.standard_normal((n, 3, 3)))
    spectra = lambda L: sym(Q @ (L[:, :, None] * np.swapaxes(Q, 1, 2)))
    e = 10.0 ** rng.uniform(-17, 0, n)
    sg = rng.choice([-1.0, 1.0], n)
    diag = np.array([np.diag(np.array(L, float)[list(perm)]) for L in ([2, -1, -1], [-2, 1, 1], [1, -1, 0], [1, 0, -1])
                     for perm in ([0, 1, 2], [1, 2, 0], [2, 0, 1], [0, 2, 1])])
    cases = {
        "random": sym(rng.standard_normal((n, 3, 3))) * (10.0 ** rng.integers(-3, 9, n))[:, None, None],
        "nearly repeated": 300.0 * spectra(np.stack([2 * sg, -sg + e, -sg - e], 1)),
        "pure-shear-like": spectra(np.stack([np.ones(n), 0.1 * e, -1 - 0.1 * e], 1)),
        "diagonal": diag,
    }
    # entries graded over twelve orders of magnitude; nearly diagonal with off-diagonals from 1e-300 to 1e-1
    grade = 10.0 ** rng.uniform(-6, 6, (n, 3))
    cases["graded"] = sym(rng.standard_normal((n, 3, 3))) * grade[:, :, None] * grade[:, None, :]
    near = np.zeros((n, 3, 3))
    near[:, [0, 1, 2], [0, 1, 2]] = rng.standard_normal((n, 3))
    for i, j in ((0, 1), (0, 2), (1, 2)):
        near[:, i, j] = near[:, j, i] = 10.0 ** rng.uniform(-300, -1, n) * rng.choice([-1.0, 1.0], n)
    cases["nearly diagonal"] = near
    for name, S in cases.items():
        S, s6 = _mandel_deviators(S)
        l, V = cport.hosford_eig(s6)
        scale = np.linalg.norm(S, axis=(1, 2))
        tol = 2e-14 if name == "nearly diagonal" else 4e-15
        assert np.max(np.linalg.norm(S @ V - V * l[:, None, :], axis=(1, 2)) / scale) < tol, name
        assert np.max(np.linalg.norm(np.swapaxes(V, 1, 2) @ V - np.eye(3), axis=(1, 2))) < 4e-15, name
        assert np.max(np.linalg.norm(np.sort(l, 1) - np.linalg.eigvalsh(S), axis=1) / scale) < 4e-15, name
    l, V = cport.hosford_eig(np.zeros((1, 6)))  # zero deviator: never a candidate point, still well defined
    assert np.array_equal(l, np.zeros((1, 3))) and np.array_equal(V[0], np.eye(3))


def test_fixed_count_root_is_accurate_for_every_exponent():
    """``q^(-1/a)`` of the criterion: Taylor start + two third-order steps, no division, no data-dependent trip count.
    Scanned over q in (0.5, 1] for every even a in [2, 64] against extended precision: a few ulp."""
    from oracle import cport

    q = np.concatenate([np.linspace(0.5, 1.0, 20001)[1:], 0.5 + 2.0 ** -np.arange(2, 53), 1.0 - 2.0 ** -np.arange(2, 54)])
    for a in range(2, 65, 2):
        w = cport.hosford_root(q, a)
        exact = q.astype(np.longdouble) ** (-np.longdouble(1.0) / a)
        assert float(np.max(np.abs(w / exact - 1.0))) < 4 * np.finfo(float).eps, a


@pytest.mark.parametrize("a", [2, 4, 6, 8, 10, 20, 64])
def test_candidate_bound_is_the_pure_shear_ratio(a):
    """sup sigma_eq / seq_Mises over all stress states = (2^(a-1)+1)^(1/a)/sqrt(3) (pure shear): the kernels finish points
    below bound * seq_Mises <= sigma_Y without an eigen-decomposition, so the bound must never be exceeded."""
    th = np.linspace(0.0, 2 * np.pi, 20001)
    s = np.stack([2.0 / 3.0 * np.cos(th - 2 * np.pi * k / 3) for k in range(3)], axis=1)  # unit von Mises stress
    v = np.concatenate([s, np.zeros_like(s)], axis=1)
    ratio = np.array([ho.sigma_eq(x, a) for x in v])
    bound = (2.0 ** (a - 1) + 1.0) ** (1.0 / a) / np.sqrt(3.0)
    assert ratio.max() <= bound * (1 + 1e-12) and ratio.max() >= bound * (1 - 1e-6)
    # a deliberately useless bound (everything is a candidate) gives the same bits as the default one
    n = 3000
    eps = synth.strain(n, 3, 8e-3, 1, 1)
    tight = ho.integrate(eps, ss.zero_state(n), dict(DEMO, a=a))
    loose = ho.integrate(eps, ss.zero_state(n), dict(DEMO, a=a, bound=1e30))
    for key in ("stress", "p", "epsp", "Ct", "flag", "n_iter", "resid", "fail"):
        assert np.array_equal(tight[key], loose[key]), key


VOCE = dict(E=70e3, nu=0.3, sig0=350.0, sigu=500.0, b=1e3, H=0.0)  # plane_elastoplasticity.py:60-69


def test_voce_hardening_exponent_two_is_the_j2_voce_oracle():
    """GeneralIsotropicHardening(elastic, Voce yield stress, Hosford a = 2) == vonMisesIsotropicHardening(elastic, Voce)."""
    n = 3000
    st_h, st_j = ss.zero_state(n), ss.zero_state(n)
    for k in range(1, 5):
        eps = synth.strain(n, 1, 1.25e-2, k, 4)
        h = ho.integrate(eps, st_h, dict(VOCE, a=2))
        j = ss.integrate(eps, st_j, VOCE)
        assert np.array_equal(h["flag"], j["flag"])
        np.testing.assert_allclose(h["stress"], j["stress"], rtol=1e-11, atol=1e-9)
        np.testing.assert_allclose(h["p"], j["p"], rtol=1e-9, atol=1e-16)
        np.testing.assert_allclose(h["Ct"], j["Ct"], rtol=1e-8, atol=1e-7 * VOCE["E"])
        st_h, st_j = ss.advance(h), ss.advance(j)
    assert h["flag"].mean() > 0.5


@pytest.mark.parametrize("a", [6, 10])
def test_voce_hardening_solution_and_tangent(a):
    props = dict(VOCE, a=a, H=25.0)
    n = 40
    (st, out) = history(props, n, 1.0e-2, 3, seed=a)[-1]
    assert out["fail"].sum() == 0 and out["flag"].sum() > 10
    lam = props["E"] * props["nu"] / (1 + props["nu"]) / (1 - 2 * props["nu"])
    mu = props["E"] / 2 / (1 + props["nu"])
    C = 2 * mu * np.eye(6)
    C[:3, :3] += lam
    sig_tr = st["stress"] + (out["strain"] - st["strain"]) @ C
    for i in np.flatnonzero(out["flag"]):
        r6, r1 = ho.implicit_residual(out["stress"][i], out["p"][i] - st["p"][i], sig_tr[i], st["p"][i], props)
        assert np.abs(r6).max() < 2e-6 * props["sig0"] and abs(r1) < 1e-9 * props["sig0"]
    eps, h = out["strain"], 1e-8
    J = np.zeros((n, 6, 6))
    for i in range(6):
        d = np.zeros(6)
        d[i] = h
        plus, minus = ho.integrate(eps + d, st, props), ho.integrate(eps - d, st, props)
        J[:, :, i] = (plus["stress"] - minus["stress"]) / (2 * h)
        J[(plus["flag"] != out["flag"]) | (minus["flag"] != out["flag"]), :, i] = np.nan
    good = ~np.isnan(J).any(axis=(1, 2))
    assert out["flag"][good].sum() > 5
    assert (np.abs(out["Ct"][good] - J[good]).max(axis=(1, 2)) / props["E"]).max() < 2e-6


def test_history_matches_reference_protocol_run():
    """tests/golden/hosford_history.npz: the reference's own Material.integrate / _vmap / DataManager drove a per-point
    Hosford material over a 3-increment history (tests/golden/make_golden.py) with the ROUND-1 restatement.  Since then
    the canonical arithmetic gained hand-placed fused multiply-adds and the a-th root a fixed-count iteration (a shorter
    dependent chain for the kernel): the batched oracle with explicit state carry must still reproduce the fixture at the
    north star's bar -- identical active sets and iteration counts, rtol 1e-10 -- in both arithmetics (pins protocol and
    regression; MFront parity itself is unpinned)."""
    import os

    from golden_check import close
    from oracle import canon

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "hosford_history.npz"))
    props = dict(zip([str(k) for k in g["props_keys"]], [float(v) for v in g["props_vals"]]))
    props["a"] = int(props["a"])
    n = g["eps1"].shape[0]
    for fused in (False, True):
        st = ss.zero_state(n)
        k = 1
        while f"eps{k}" in g:
            if fused:
                out = ho.integrate(g[f"eps{k}"], st, props)
            else:
                with canon.unfused():
                    out = ho.integrate(g[f"eps{k}"], st, props)
            # active set of the fixture: the cumulated plastic strain grew
            p_prev = st["p"].reshape(n)
            assert np.array_equal(out["flag"].astype(bool), g[f"isv{k}"][:, 0] > p_prev)
            close(out["stress"], g[f"flux{k}"], "stress")
            close(out["p"], g[f"isv{k}"][:, 0], "p")
            close(out["epsp"], g[f"isv{k}"][:, 1:], "epsp")
            close(out["Ct"], g[f"Ct{k}"], "Ct")
            st = ss.advance(out)
            k += 1
        assert k == 4 and out["flag"].any() and not out["flag"].all()
