"""Canonical IEEE-754 building blocks shared (by construction, not by import) with the CUDA kernels.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

``exp_c`` is a hand-written double-precision exponential made only of correctly rounded
operations (``*``, ``+``, ``-``, ``rint``, ``ldexp``) in a fixed order, with no fused
multiply-add.  numpy evaluates every ufunc separately (never contracting ``a*b+c``), the C
oracle is built with ``-ffp-contract=off`` and the CUDA kernels with ``-fmad=false``; all three
therefore produce the same bits.  Its accuracy (< 1.5 ulp against libm on the range the Voce law
``sigma_Y(p) = sig0 + (sigu-sig0)(1-exp(-b p))`` uses, reference ``tests/test_FeFp_jax.py:14-15``)
is checked in ``tests/test_oracle_canon.py``.
"""

import numpy as np

LOG2E = 1.4426950408889634  # 0x3FF71547652B82FE
LN2_HI = 6.93147180369123816490e-01  # 0x3FE62E42FEE00000 (low 21 bits zero: k*LN2_HI is exact)
LN2_LO = 1.90821492927058770002e-10  # 0x3DEA39EF35793C76
EXP_CLAMP = 700.0

# 1/n!, n = 13 .. 0 (Horner order).  |r| <= ln2/2 => truncation error r^14/14! < 5e-18.
EXP_POLY = (
    1.0 / 6227020800.0,
    1.0 / 479001600.0,
    1.0 / 39916800.0,
    1.0 / 3628800.0,
    1.0 / 362880.0,
    1.0 / 40320.0,
    1.0 / 5040.0,
    1.0 / 720.0,
    1.0 / 120.0,
    1.0 / 24.0,
    1.0 / 6.0,
    0.5,
    1.0,
    1.0,
)


def exp_c(x):
    """exp(x) with the canonical operation order; vectorised over numpy arrays."""
    x = np.asarray(x, dtype=np.float64)
    inr = (x >= -EXP_CLAMP) & (x <= EXP_CLAMP)
    xs = np.where(inr, x, 0.0)
    k = np.rint(xs * LOG2E)
    r = (xs - k * LN2_HI) - k * LN2_LO
    y = np.full_like(xs, EXP_POLY[0])
    for c in EXP_POLY[1:]:
        y = y * r + c
    y = np.ldexp(y, k.astype(np.int32))
    y = np.where(x < -EXP_CLAMP, 0.0, y)
    y = np.where(x > EXP_CLAMP, np.inf, y)
    y = np.where(np.isnan(x), np.nan, y)
    return y


def lame(E, nu):
    """(lambda, mu) with the reference's operation order (``python_materials/elasticity.py:12-13``)."""
    E = np.asarray(E, dtype=np.float64)
    nu = np.asarray(nu, dtype=np.float64)
    lam = E * nu / (1 + nu) / (1 - 2 * nu)
    mu = E / 2 / (1 + nu)
    return lam, mu
