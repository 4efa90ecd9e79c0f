"""The MFront behaviour ``IsotropicPlasticHosfordFlowLinear`` solved in 40-digit arithmetic.  TEST INFRASTRUCTURE ONLY.

A third, independent statement of ``demos/multimaterials/IsotropicPlasticHosfordFlowLinear.mfront:1-27`` -- the
``StandardElastoViscoPlasticity`` brick with a Hooke stress potential, the Hosford criterion and linear (optionally Voce)
isotropic hardening, implicit scheme with theta = 1 -- written the way MFront's ``Implicit`` DSL poses it: seven unknowns
(the increment of the elastic strain, 6 Mandel components, and the increment of the equivalent plastic strain) with

    f_eel = d_eel - d_eto + dp n(sigma),        f_p = (sigma_eq(sigma) - R(p_old + dp)) / E,
    sigma = C : (eel_old + d_eel),              n = d sigma_eq / d sigma,

solved by ``mpmath.findroot`` at 40 digits; the flow direction comes from the spectral decomposition
(``mpmath.eigsy``) and the analytic derivative of the criterion with respect to the principal stresses -- no divided
differences, no scaling tricks, no shared code with ``oracle/c/dxm_oracle_hosford.c``.  The consistent tangent is the
central difference of that solution map (step 1e-18 relative: truncation and rounding both below 1e-25).

Used by ``tests/test_oracle_hosford_mp.py`` to bound the distance between the canonical double-precision arithmetic
(which the CUDA kernels reproduce bit for bit) and the exact solution of the equations MFront solves.  Parity with
MFront's own floating-point results stays unpinned (TFEL / MGIS absent): this pins the MATHEMATICS.
"""

import mpmath as mp

mp.mp.dps = 40
R2 = mp.sqrt(2)


def _tensor(v):
    return mp.matrix([[v[0], v[3] / R2, v[4] / R2], [v[3] / R2, v[1], v[5] / R2], [v[4] / R2, v[5] / R2, v[2]]])


def _mandel(T):
    return [T[0, 0], T[1, 1], T[2, 2], R2 * T[0, 1], R2 * T[0, 2], R2 * T[1, 2]]


def sigma_eq_and_normal(sig, a):
    """Hosford equivalent stress of a Mandel 6-vector and its gradient (Mandel 6-vector)."""
    s, Q = mp.eigsy(_tensor(sig))
    d = [s[0] - s[1], s[1] - s[2], s[2] - s[0]]
    q = (d[0] ** a + d[1] ** a + d[2] ** a) / 2  # a even: |.|^a == (.)^a
    phi = q ** (mp.mpf(1) / a)
    if phi == 0:
        return phi, [mp.mpf(0)] * 6
    # d phi / d s_i = phi^(1-a) / 2 * (sum over the differences that contain s_i, with sign) of d^(a-1)
    c = phi ** (1 - a) / 2
    g = [c * (d[0] ** (a - 1) - d[2] ** (a - 1)), c * (d[1] ** (a - 1) - d[0] ** (a - 1)),
         c * (d[2] ** (a - 1) - d[1] ** (a - 1))]
    N = mp.zeros(3, 3)
    for i in range(3):
        v = Q[:, i]
        N += g[i] * (v * v.T)
    return phi, _mandel(N)


def yield_radius(p, props):
    r = mp.mpf(props["sig0"]) + mp.mpf(props.get("H", 0.0)) * p
    if props.get("sigu") is not None and props.get("b"):
        r += (mp.mpf(props["sigu"]) - mp.mpf(props["sig0"])) * (1 - mp.exp(-mp.mpf(props["b"]) * p))
    return r


def _hooke(props):
    E, nu = mp.mpf(props["E"]), mp.mpf(props["nu"])
    lam, mu = E * nu / (1 + nu) / (1 - 2 * nu), E / 2 / (1 + nu)
    return lam, mu


def _stress(eel, lam, mu):
    tr = eel[0] + eel[1] + eel[2]
    return [2 * mu * eel[i] + (lam * tr if i < 3 else 0) for i in range(6)]


def integrate_point(eps, eps_old, epsp_old, p_old, props, start=None):
    """One Gauss point.  Arguments: Mandel 6-vectors / scalars (floats or mpf).  ``start``: optional (d_eel, dp) initial
    guess (e.g. the double-precision solution -- the 40-digit root is unique, the guess only saves iterations).
    Returns ``dict(stress, p, epsp, plastic)`` in mpf."""
    a = int(props["a"])
    lam, mu = _hooke(props)
    eps = [mp.mpf(x) for x in eps]
    eel_old = [mp.mpf(x) - mp.mpf(y) for x, y in zip(eps_old, epsp_old)]
    deto = [x - mp.mpf(y) for x, y in zip(eps, eps_old)]
    p_old = mp.mpf(p_old)
    eel_tr = [x + y for x, y in zip(eel_old, deto)]
    sig_tr = _stress(eel_tr, lam, mu)
    phi_tr, _ = sigma_eq_and_normal(sig_tr, a)
    if phi_tr <= yield_radius(p_old, props):
        return dict(stress=sig_tr, p=p_old, epsp=[mp.mpf(x) for x in epsp_old], plastic=False)
    E = mp.mpf(props["E"])

    def residual(*x):
        deel, dp = list(x[:6]), x[6]
        sig = _stress([u + v for u, v in zip(eel_old, deel)], lam, mu)
        phi, n = sigma_eq_and_normal(sig, a)
        return [deel[i] - deto[i] + dp * n[i] for i in range(6)] + [(phi - yield_radius(p_old + dp, props)) / E]

    x0 = list(start[0]) + [start[1]] if start is not None else deto + [mp.mpf(0)]
    x = mp.findroot(residual, [mp.mpf(v) for v in x0], tol=mp.mpf(10) ** -60, maxsteps=60, verify=False)
    res = residual(*x)
    assert max(abs(r) for r in res) < mp.mpf(10) ** -30, "40-digit solve did not converge"
    deel, dp = [x[i] for i in range(6)], x[6]
    eel = [u + v for u, v in zip(eel_old, deel)]
    return dict(stress=_stress(eel, lam, mu), p=p_old + dp,
                epsp=[mp.mpf(e) + (d - de) for e, d, de in zip(epsp_old, deto, deel)], plastic=True,
                start=(deel, dp))


def tangent_point(eps, eps_old, epsp_old, p_old, props, start=None, h=None):
    """d stress / d strain (6x6, Mandel) of the solution map by central differences in 40-digit arithmetic."""
    scale = max(abs(mp.mpf(x)) for x in eps) or mp.mpf(1)
    h = mp.mpf(h) if h is not None else scale * mp.mpf(10) ** -18
    base = integrate_point(eps, eps_old, epsp_old, p_old, props, start)
    st = base.get("start")
    Ct = mp.zeros(6, 6)
    for j in range(6):
        ep, em = [mp.mpf(x) for x in eps], [mp.mpf(x) for x in eps]
        ep[j] += h
        em[j] -= h
        sp = integrate_point(ep, eps_old, epsp_old, p_old, props, st)["stress"]
        sm = integrate_point(em, eps_old, epsp_old, p_old, props, st)["stress"]
        for i in range(6):
            Ct[i, j] = (sp[i] - sm[i]) / (2 * h)
    return Ct
