"""Small-strain Hosford plasticity oracle.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

``integrate`` is the canonical-arithmetic restatement (plain C, ``oracle/c/dxm_oracle_hosford.c``, through
``oracle/cport.py``) of the MFront behaviour ``demos/multimaterials/IsotropicPlasticHosfordFlowLinear.mfront:1-27``
(Hooke + Hosford criterion ``{a: 10}`` + linear isotropic hardening, implicit, theta = 1) that the reference's
multi-material demo puts in the matrix phase (``demos/multimaterials/multimaterials.py:245-254``).  **Parity with
MFront is unpinned**: TFEL/MFront/MGIS are not in this image and the generated code is not in the tree.

The functions below the line are an independent, deliberately naive numpy statement of the same equations (eigvalsh,
finite-difference flow direction) that the tests use to check the canonical restatement without sharing code with it.
"""

import numpy as np

from . import cport
from .small_strain import advance, zero_state  # noqa: F401  (same state layout: strain, stress, p, epsp)


def integrate(eps, state, props, newton_cap=25, rtol=1e-12):
    """props: ``E, nu, sig0`` (R0), ``H`` and optionally ``sigu, b`` (Voce term) -- scalars or per-point arrays -- and the
    even integer exponent ``a``.
    Returns the same dictionary as ``oracle.small_strain.integrate``."""
    return cport.hosford(eps, state, dict(props, H=props.get("H", 0.0)), newton_cap, rtol)


# ---------------------------------------------------------------------------------------------------------------
R2 = np.sqrt(2.0)


def mandel_to_tensor(v):
    return np.array([[v[0], v[3] / R2, v[4] / R2], [v[3] / R2, v[1], v[5] / R2], [v[4] / R2, v[5] / R2, v[2]]])


def sigma_eq(v, a):
    """Hosford equivalent stress of a Mandel 6-vector, straight from the definition."""
    s = np.linalg.eigvalsh(mandel_to_tensor(v))
    return (0.5 * (abs(s[0] - s[1]) ** a + abs(s[1] - s[2]) ** a + abs(s[2] - s[0]) ** a)) ** (1.0 / a)


def flow_direction(v, a, h=None):
    """d sigma_eq / d sigma (Mandel) by central differences."""
    v = np.asarray(v, dtype=float)
    h = h or 1e-6 * max(1.0, np.abs(v).max())
    n = np.zeros(6)
    for i in range(6):
        e = np.zeros(6)
        e[i] = h
        n[i] = (sigma_eq(v + e, a) - sigma_eq(v - e, a)) / (2 * h)
    return n


def implicit_residual(sig, dp, sig_tr, p_old, props):
    """Residual of the backward-Euler system in stress form: sigma - sigma_tr + 2 mu dp n(sigma) (n is deviatoric, so
    C : n = 2 mu n) and the yield condition; returns (6-vector, scalar)."""
    mu = props["E"] / 2 / (1 + props["nu"])
    n = flow_direction(sig, props["a"])
    return sig - sig_tr + 2 * mu * dp * n, sigma_eq(sig, props["a"]) - yield_stress(p_old + dp, props)


def yield_stress(p, props):
    """sig0 + H p + (sigu - sig0)(1 - exp(-b p)) -- linear for sigu = sig0, Voce for H = 0 (tests/test_FeFp_jax.py:14-15)."""
    sig0 = props["sig0"]
    return sig0 + props.get("H", 0.0) * p + (props.get("sigu", sig0) - sig0) * (1.0 - np.exp(-props.get("b", 0.0) * p))
