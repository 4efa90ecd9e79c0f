"""Canonical IEEE-754 building blocks shared (by construction, not by import) with the CUDA kernels.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

The canonical arithmetic is a fixed sequence of correctly rounded IEEE-754 operations: ``+ - * / sqrt rint ldexp``
and the EXPLICIT fused multiply-add :func:`fma` / :func:`fms` / :func:`fnma`, placed by hand at the same positions
in the three implementations: here, in the plain-C oracle (``oracle/c/dxm_canon.h``, built with
``-ffp-contract=off``) and in the CUDA kernels (``csrc/dxm_canon.cuh``, built with ``-fmad=false``) -- no compiler
ever fuses (or splits) on its own, so all three produce the same bits while the kernels issue one DFMA where the
round-1 arithmetic needed DMUL + DADD.  numpy has no fma ufunc: the primitive is C99 ``fma()`` applied element-wise
(``oracle/c/dxm_fma_vec.c``), checked against exact rational arithmetic in ``tests/test_oracle_canon.py``.
``FUSED = False`` (or :func:`unfused`) splits every fma into two roundings again, i.e. the round-1 arithmetic.

``exp_c`` is a hand-written double-precision exponential (Cody-Waite reduction, degree-13 Horner in fused steps,
exact scaling).  Its accuracy (< 1.5 ulp against libm on the range the Voce law
``sigma_Y(p) = sig0 + (sigu-sig0)(1-exp(-b p))`` uses, reference ``tests/test_FeFp_jax.py:14-15``)
is checked in ``tests/test_oracle_canon.py``.
"""

import contextlib
import ctypes

import numpy as np

FUSED = True


@contextlib.contextmanager
def unfused():
    """Evaluate the oracles with every explicit fma split into ``*`` then ``+`` (the round-1 arithmetic)."""
    global FUSED
    old, FUSED = FUSED, False
    try:
        yield
    finally:
        FUSED = old


def _fma_lib(a, b, c):
    from . import cport

    lib = cport.load()
    a, b, c = (np.asarray(v, dtype=np.float64) for v in (a, b, c))
    shape = np.broadcast_shapes(a.shape, b.shape, c.shape)
    ops, strides = [], []
    for v in (a, b, c):
        if v.ndim == 0:
            ops.append(np.ascontiguousarray(v.reshape(1)))
            strides.append(0)
        else:
            ops.append(np.ascontiguousarray(np.broadcast_to(v, shape)).reshape(-1))
            strides.append(1)
    out = np.empty(shape, dtype=np.float64)
    lib.dxo_fma_vec(ctypes.c_int64(out.size), *[x for o, st in zip(ops, strides)
                                                for x in (ctypes.c_void_p(o.ctypes.data), ctypes.c_int64(st))],
                    ctypes.c_void_p(out.ctypes.data))
    return out if out.ndim else float(out)


def fma(a, b, c):
    """``a*b + c`` with ONE rounding (two when ``FUSED`` is off); broadcasts like numpy."""
    if not FUSED:
        return a * b + c
    return _fma_lib(a, b, c)


def fms(a, b, c):
    """``a*b - c`` with one rounding."""
    if not FUSED:
        return a * b - c
    return _fma_lib(a, b, -np.asarray(c, dtype=np.float64))


def fnma(a, b, c):
    """``c - a*b`` with one rounding."""
    if not FUSED:
        return c - a * b
    return _fma_lib(-np.asarray(a, dtype=np.float64), b, c)


def dot3(a0, b0, a1, b1, a2, b2):
    """``(a0 b0 + a1 b1) + a2 b2`` as one product and two fused steps."""
    return fma(a2, b2, fma(a1, b1, a0 * b0))


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
    r = fnma(k, LN2_LO, fnma(k, LN2_HI, xs))
    y = np.full_like(xs, EXP_POLY[0])
    for c in EXP_POLY[1:]:
        y = fma(y, r, c)
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
