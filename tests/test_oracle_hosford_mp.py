"""The Hosford oracle (canonical double-precision arithmetic, reproduced bit for bit by the CUDA kernel) against the
SAME equations solved in 40-digit arithmetic the way MFront's ``Implicit`` DSL poses them
(``demos/multimaterials/IsotropicPlasticHosfordFlowLinear.mfront:1-27``; ``oracle/hosford_mp.py``: seven unknowns,
``mpmath.findroot``, spectral flow direction, no shared code).  Bounds the distance between what the GPU computes and the
exact solution of the reference behaviour's equations: a few 1e-13 relative for stress and state, 1e-12 for the
consistent tangent -- three orders of magnitude inside the north star's rtol 1e-10.  MFront's own floating-point results
stay unpinned (TFEL / MGIS absent); this pins the mathematics.  CPU only."""

import numpy as np
import pytest

mp = pytest.importorskip("mpmath")

from oracle import hosford as ho  # noqa: E402
from oracle import hosford_mp as hm  # noqa: E402
from oracle import small_strain as ss  # noqa: E402
from oracle import synth  # noqa: E402

DEMO = dict(E=70e3, nu=0.3, sig0=200.0, H=10.0, a=10)  # multimaterials.py:245-254


def _f(v):
    return np.array([float(x) for x in v])


def _start(eps, st, out, i):
    d_eel = (eps[i] - st["strain"][i]) - (out["epsp"][i] - st["epsp"][i])
    return list(d_eel), out["p"][i] - st["p"][i]


CASES = [dict(DEMO), dict(DEMO, a=6), dict(DEMO, a=20, H=2e3),
         dict(E=70e3, nu=0.3, sig0=350.0, sigu=500.0, b=1e3, H=0.0, a=8)]


@pytest.mark.parametrize("props", CASES, ids=lambda p: f"a{p['a']}" + ("-voce" if "sigu" in p else ""))
def test_stress_and_state_vs_40_digit_solution(props):
    n = 16
    st = ss.zero_state(n)
    plastic = 0
    for k in (1, 2):
        eps = synth.strain(n, props["a"], 1.0e-2, k, 2)
        if k == 1:
            eps[0, 1:] = 0.0  # uniaxial strain: a repeated eigenvalue
            eps[1] = [3e-3, 3e-3, -6e-3, 0, 0, 0]  # axisymmetric
        out = ho.integrate(eps, st, props)
        assert out["fail"].sum() == 0
        for i in range(n):
            ref = hm.integrate_point(eps[i], st["strain"][i], st["epsp"][i], st["p"][i], props,
                                     start=_start(eps, st, out, i) if out["flag"][i] else None)
            assert ref["plastic"] == bool(out["flag"][i])
            s = _f(ref["stress"])
            scale = np.abs(s).max()
            assert np.abs(out["stress"][i] - s).max() <= 2e-12 * scale, (i, k)
            assert abs(out["p"][i] - float(ref["p"])) <= 2e-12 * max(float(ref["p"]), 1e-3), (i, k)
            assert np.abs(out["epsp"][i] - _f(ref["epsp"])).max() <= 2e-12 * max(np.abs(eps[i]).max(), 1e-3), (i, k)
            plastic += int(ref["plastic"])
        st = ss.advance(out)
    assert plastic >= n


def test_consistent_tangent_vs_40_digit_derivative():
    n = 6
    worst = 0.0
    for props in (dict(DEMO), CASES[3]):
        st = ss.zero_state(n)
        eps = synth.strain(n, 1, 1.0e-2, 1, 1)
        out = ho.integrate(eps, st, props)
        pts = [i for i in range(n) if out["flag"][i]][:2]
        assert len(pts) == 2
        for i in pts:
            Ct = hm.tangent_point(eps[i], st["strain"][i], st["epsp"][i], st["p"][i], props, start=_start(eps, st, out, i))
            C = np.array([[float(Ct[r, c]) for c in range(6)] for r in range(6)])
            err = np.abs(C - out["Ct"][i]).max() / np.abs(C).max()
            worst = max(worst, err)
            assert err <= 1e-11, (props["a"], i, err)
            assert np.abs(C - C.T).max() <= 1e-20 * np.abs(C).max()  # associated flow: the exact tangent is symmetric
    assert worst > 0.0


def test_extreme_regimes_vs_40_digit_solution():
    """Steps of 35 x the yield strain, the largest supported exponent, nearly incompressible elasticity with stiff hardening."""
    for props, amp in ((dict(DEMO), 1e-1), (dict(DEMO, a=64), 1e-2), (dict(E=210e3, nu=0.45, sig0=50.0, H=1e4, a=12), 3e-2)):
        n = 5
        st = ss.zero_state(n)
        eps = synth.strain(n, 7, amp, 1, 1)
        out = ho.integrate(eps, st, props)
        assert out["fail"].sum() == 0 and out["flag"].all()
        for i in range(n):
            ref = hm.integrate_point(eps[i], st["strain"][i], st["epsp"][i], st["p"][i], props, start=_start(eps, st, out, i))
            s = _f(ref["stress"])
            assert np.abs(out["stress"][i] - s).max() <= 3e-12 * np.abs(s).max(), (props["a"], i)
            assert abs(out["p"][i] - float(ref["p"])) <= 1e-11 * float(ref["p"]), (props["a"], i)
