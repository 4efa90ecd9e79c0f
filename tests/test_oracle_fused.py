"""The hand-placed fused multiply-adds of the canonical arithmetic (CPU only).

* the fma primitive of the numpy oracle is the correctly rounded ``a*b + c`` (checked against exact rationals);
* numpy oracle == plain-C oracle bit for bit in BOTH arithmetics (fused, and un-fused = round 1);
* fused vs un-fused (today's vs round 1's oracle): identical active sets and local iteration counts, stress / state /
  tangent within the north star's rtol 1e-10 -- over seeded histories of all three kernel families.
"""
from fractions import Fraction

import numpy as np
import pytest
from golden_check import close, same_active_set

from oracle import canon, cport, fefp, hosford, synth
from oracle import small_strain as ss


def test_fma_is_correctly_rounded():
    rng = np.random.default_rng(1)
    a, b = rng.standard_normal(500), rng.standard_normal(500)
    c = -(a * b) * (1.0 + 1e-13 * rng.standard_normal(500))  # heavy cancellation: the case where fusing matters
    for f, sign_ab, sign_c in ((canon.fma, 1, 1), (canon.fms, 1, -1), (canon.fnma, -1, 1)):
        out = f(a, b, c)
        exact = [float(sign_ab * Fraction(x) * Fraction(y) + sign_c * Fraction(z)) for x, y, z in zip(a, b, c)]
        assert np.array_equal(out, np.array(exact))
    assert np.count_nonzero(canon.fma(a, b, c) != a * b + c) > 400  # and it is not the two-rounding result
    assert canon.fma(2.0, 3.0, 1.0) == 7.0  # scalars
    assert np.array_equal(canon.fma(a, 2.0, 1.0), 2.0 * a + 1.0)  # broadcast; exact product -> same bits
    with canon.unfused():
        assert np.array_equal(canon.fma(a, b, c), a * b + c)
    assert canon.FUSED


J2 = dict(E=70e3, nu=0.3, sig0=350.0, sigu=500.0, b=1e3)
FE = dict(E=70e3, nu=0.3, sig0=500.0, sigu=750.0, b=1000.0)
HO = dict(E=70e3, nu=0.3, sig0=200.0, H=10.0, a=10)


def _histories(kind, n):
    if kind == "j2":
        return [synth.strain(n, 0, 1.25e-2, k, 4) for k in range(1, 5)], ss.zero_state(n), J2
    if kind == "hosford":
        return [synth.strain(n, 3, 1.25e-2, k, 3) for k in range(1, 4)], ss.zero_state(n), HO
    return [synth.defgrad(n, 0, 3e-2, k, 3) for k in range(1, 4)], fefp.virgin_state(n), FE


RUN = {"j2": (ss.integrate, cport.small_strain, ss.advance, ("stress", "p", "epsp", "Ct")),
       "hosford": (None, cport.hosford, ss.advance, ("stress", "p", "epsp", "Ct")),
       "fefp": (fefp.integrate, cport.fefp, fefp.advance, ("PK1", "p", "be_bar", "Ct"))}


@pytest.mark.parametrize("kind", ["j2", "hosford", "fefp"])
def test_fused_vs_round1_arithmetic(kind):
    n = 20000
    grads, st, props = _histories(kind, n)
    py, c, adv, fields = RUN[kind]
    st_u = st
    worst = 0.0
    for g in grads:
        out = c(g, st, props)
        with canon.unfused():
            ref = c(g, st_u, props)
        if py is not None:  # the numpy restatement is the same arithmetic, bit for bit, in both modes
            out_py = py(g, st, props)
            with canon.unfused():
                ref_py = py(g, st_u, props)
            for k in fields + ("flag", "n_iter", "resid", "fail"):
                assert np.array_equal(out_py[k], out[k]), k
                assert np.array_equal(ref_py[k], ref[k]), k
        same_active_set(out, ref, borderline=1e-4 if kind == "hosford" else 0.0)
        for k in fields:
            close(out[k], ref[k], k)
            worst = max(worst, float(np.max(np.abs(out[k] - ref[k])) / np.max(np.abs(ref[k]))))
        assert any(np.count_nonzero(out[k] != ref[k]) for k in fields)  # the two arithmetics do differ in the last bits
        st, st_u = adv(out), adv(ref)
    assert out["flag"].mean() > 0.5 and out["fail"].sum() == 0
    assert worst < 1e-12  # in fact a few 1e-14: two orders inside the tolerance
