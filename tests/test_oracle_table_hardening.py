"""Piecewise-linear (tabulated) isotropic hardening in the small-strain oracle -- SURVEY 8(f) rank 4, the device-side
stand-in for the arbitrary ``yield_stress`` callable of jaxmat's ``vonMisesIsotropicHardening``.  Checks that need no
reference: a two-point table is the linear-hardening closed form bit for bit (itself pinned to the MFront source),
a dense sampling of the Voce law converges to the Voce update, yield consistency on the table, exact tangent."""
import numpy as np

from oracle import small_strain as ss
from oracle import synth


def history(props, n, amp, K):
    st = ss.zero_state(n)
    outs = []
    for k in range(1, K + 1):
        out = ss.integrate(synth.strain(n, 0, amp, k, K), st, props)
        outs.append((st, out))
        st = ss.advance(out)
    return outs


def test_two_point_table_is_linear_hardening_bitwise():
    n = 1500
    lin = dict(E=70e3, nu=0.3, sig0=250.0, H=5e3)
    tab = dict(E=70e3, nu=0.3, table=([0.0, 1.0], [250.0, 5250.0]))
    for (_, a), (_, b) in zip(history(lin, n, 1.25e-2, 3), history(tab, n, 1.25e-2, 3)):
        for key in ("stress", "p", "epsp", "Ct", "flag"):
            assert np.array_equal(a[key], b[key]), key
        assert b["n_iter"].max() == 0 and a["flag"].any()


def test_dense_table_converges_to_voce_and_is_consistent():
    n = 2000
    voce = dict(E=70e3, nu=0.3, sig0=350.0, sigu=500.0, b=1e3)
    law = lambda p: 350.0 + 150.0 * (1 - np.exp(-1e3 * p))  # noqa: E731
    errs = []
    for npts in (16, 64):
        pk = np.concatenate([[0.0], np.geomspace(2e-6, 0.2, npts - 1)])
        tab = dict(E=70e3, nu=0.3, table=(pk, law(pk)))
        a = history(voce, n, 1.25e-2, 4)[-1][1]
        st, b = history(tab, n, 1.25e-2, 4)[-1]
        assert np.array_equal(a["flag"], b["flag"])
        errs.append(np.abs(a["stress"] - b["stress"]).max() / np.abs(a["stress"]).max())
        # yield consistency on the piecewise-linear curve, dp >= 0, crossings happen
        pl = b["flag"] == 1
        sig = b["stress"]
        s = sig.copy()
        s[:, :3] -= sig[:, :3].mean(1, keepdims=True)
        seq = np.sqrt(1.5 * (s * s).sum(1))
        sy = np.interp(b["p"], pk, law(pk))
        assert np.abs(seq - sy)[pl].max() < 1e-9 * 350 and (b["p"] - st["p"].ravel()).min() >= 0
        assert b["n_iter"].max() >= 1 and b["fail"].sum() == 0
    assert errs[1] < errs[0] / 4 and errs[1] < 2e-4


def test_table_tangent_matches_finite_differences_and_symmetry():
    n = 600
    pk = np.array([0.0, 1e-3, 4e-3, 2e-2, 0.2])
    tab = dict(E=70e3, nu=0.3, table=(pk, np.array([300.0, 380.0, 430.0, 470.0, 520.0])))
    st, out = history(tab, n, 1.25e-2, 3)[-1]
    eps = synth.strain(n, 0, 1.25e-2, 3, 3)
    Ct = out["Ct"]
    assert np.array_equal(Ct, Ct.transpose(0, 2, 1)) and 0.3 < out["flag"].mean() < 0.95
    h = 1e-7
    fd = np.zeros_like(Ct)
    for i in range(6):
        ep, em = eps.copy(), eps.copy()
        ep[:, i] += h
        em[:, i] -= h
        op, om = ss.integrate(ep, st, tab), ss.integrate(em, st, tab)
        fd[:, :, i] = (op["stress"] - om["stress"]) / (2 * h)
    # points whose perturbed solutions stay on the same segment (the curve has kinks at the table points)
    same = np.ones(n, dtype=bool)
    for i in range(6):
        ep = eps.copy()
        ep[:, i] += h
        same &= ss.integrate(ep, st, tab)["n_iter"] == out["n_iter"]
    err = np.abs(fd - Ct).max(axis=(1, 2)) / np.abs(Ct).max(axis=(1, 2))
    assert same.mean() > 0.9 and err[same].max() < 1e-6
