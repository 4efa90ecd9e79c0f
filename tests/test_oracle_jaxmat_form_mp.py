"""The canonical double-precision oracle of the two jaxmat behaviours (what the CUDA kernels compute bit for bit) against
jaxmat's OWN branch-free equations solved in 40-digit arithmetic (``oracle/jaxmat_form_mp.py``: Fischer-Burmeister scalar
problem for ``vonMisesIsotropicHardening``, the seven-unknown system for ``FeFpJ2Plasticity``, tangents as derivatives of
the solution map).  ``tests/test_oracle_jaxmat_form.py`` holds the oracle to the same formulation in double precision at
rtol 1e-10 over the full golden histories; this test says how far the oracle is from the EXACT solution: stress within 1e-12
relative, the plastic multiplier within the local Newton tolerance (a few 1e-12), tangents within 1e-11 -- two to four orders inside the north star's rtol 1e-10.  jaxmat's own floating-point
output is still what would close the pin (jax / jaxmat are not installable here).  CPU only."""

import numpy as np
import pytest

mp = pytest.importorskip("mpmath")

from oracle import fefp  # noqa: E402
from oracle import jaxmat_form_mp as jmp  # noqa: E402
from oracle import small_strain as ss  # noqa: E402
from oracle import synth  # noqa: E402

VOCE = dict(E=70e3, nu=0.3, sig0=350.0, sigu=500.0, b=1e3)  # plane_elastoplasticity.py:60-69
FEFP = dict(E=70e3, nu=0.3, sig0=500.0, sigu=750.0, b=1000.0)  # tests/test_FeFp_jax.py:6-19


def _f(v):
    return np.array([float(x) for x in v])


def test_small_strain_voce_vs_40_digit_solution():
    n = 64
    st = ss.zero_state(n)
    plastic = 0
    for k in (1, 2, 3, 4):
        eps = synth.strain(n, 0, 1.25e-2, k, 4)
        out = ss.integrate(eps, st, VOCE)
        for i in range(n):
            ref = jmp.j2_point(eps[i], st["strain"][i], st["stress"][i], st["p"][i], VOCE, dp0=float(out["p"][i] - st["p"][i]))
            s = _f(ref["stress"])
            assert (ref["dp"] > 0) == bool(out["flag"][i])
            assert np.abs(out["stress"][i] - s).max() <= 2e-12 * np.abs(s).max()
            assert abs(out["p"][i] - float(ref["p"])) <= 1e-11 * float(ref["p"]) + 1e-18  # the local Newton stops at 1e-12 seq on the residual
            assert np.abs(out["epsp"][i] - st["epsp"][i] - _f(ref["depsp"])).max() <= 1e-11 * max(np.abs(_f(ref["depsp"])).max(), 1e-6)
            plastic += int(out["flag"][i])
        for i in [j for j in range(n) if out["flag"][j]][:3] + [j for j in range(n) if not out["flag"][j]][:1]:
            Ct = jmp.j2_tangent(eps[i], st["strain"][i], st["stress"][i], st["p"][i], VOCE, dp0=float(out["p"][i] - st["p"][i]))
            C = np.array([[float(Ct[r, c]) for c in range(6)] for r in range(6)])
            assert np.abs(C - out["Ct"][i]).max() <= 1e-12 * np.abs(C).max()
        st = ss.advance(out)
    assert plastic > n


def _fefp_check(F, st, out, i, tangent=False):
    start = (float(out["p"][i] - st["p"][i]), out["be_bar"][i])
    ref = jmp.fefp_point(F[i], st["F"][i], st["be_bar"][i], st["p"][i], FEFP, start=start)
    P = _f(ref["PK1"])
    assert (ref["dp"] > 1e-30) == bool(out["flag"][i])
    assert np.abs(out["PK1"][i] - P).max() <= 5e-12 * np.abs(P).max()
    assert abs(out["p"][i] - float(ref["p"])) <= 1e-12
    assert np.abs(out["be_bar"][i] - _f(ref["be_bar"])).max() <= 1e-12
    if tangent:
        Ct = jmp.fefp_tangent(F[i], st["F"][i], st["be_bar"][i], st["p"][i], FEFP, start=start)
        C = np.array([[float(Ct[r, c]) for c in range(9)] for r in range(9)])
        assert np.abs(C - out["Ct"][i].reshape(9, 9)).max() <= 1e-11 * np.abs(C).max()


def test_finite_strain_vs_40_digit_solution():
    n = 16
    st = fefp.virgin_state(n)
    plastic = 0
    for k in (1, 2):
        F = synth.defgrad(n, 0, 3e-2, k, 2)
        out = fefp.integrate(F, st, FEFP)
        assert out["fail"].sum() == 0
        pl = [j for j in range(n) if out["flag"][j]]
        for i in range(n):
            _fefp_check(F, st, out, i, tangent=(i in pl[:2]) or (k == 1 and i == [j for j in range(n) if not out["flag"][j]][:1]))
        plastic += len(pl)
        st = fefp.advance(out)
    assert plastic >= n


def test_reference_fefp_script_vs_40_digit_solution():
    """The loading of ``tests/test_FeFp_jax.py:21-33`` (isochoric-like stretch to 2 %, 19 steps), one point."""
    nb, eps, nsteps = 1, 2e-2, 20
    st = fefp.virgin_state(nb)
    seen = 0
    for t in np.linspace(0, 1.0, nsteps)[1:]:
        F = np.zeros((nb, 9))
        F[:, 0] = 1 + eps * t
        F[:, [1, 2]] = 1 - eps / 2 * t
        out = fefp.integrate(F, st, FEFP)
        _fefp_check(F, st, out, 0, tangent=t > 0.9)
        seen += int(out["flag"][0])
        st = fefp.advance(out)
    assert seen > 5


def test_finite_strain_large_deformations_vs_40_digit_solution():
    """|F - I| up to 0.3 per increment (every point plastic): the reduced 2x2 solve of the canonical oracle stays within
    5e-12 of the exact solution of jaxmat's seven-unknown system."""
    for amp in (0.1, 0.3):
        n = 6
        st = fefp.virgin_state(n)
        for k in (1, 2):
            F = synth.defgrad(n, 3, amp, k, 2)
            out = fefp.integrate(F, st, FEFP)
            assert out["fail"].sum() == 0 and out["flag"].all()
            for i in range(n):
                _fefp_check(F, st, out, i)
            st = fefp.advance(out)
