"""jaxmat's branch-free local problems (``oracle/jaxmat_form.py`` has them in double precision with complex-step
derivatives) solved in 40-digit arithmetic, point by point.  TEST INFRASTRUCTURE ONLY.

* ``vonMisesIsotropicHardening``: ``FB(-f/E, dp) = 0``, ``f = seq_el - 3 mu dp - sigma_Y(p_old + dp)``;
* ``FeFpJ2Plasticity``: seven unknowns ``(dp, be_bar)`` with ``FB(-f_y/E, dp) = 0`` and
  ``dev(be - be_tr) + 2/3 dp tr(be) n + 1 (det be - 1) = 0``, ``tau = mu dev(be) + kappa/2 (J^2 - 1) 1``, ``P = tau F^-T``;

(call sites ``demos/jax/elastoplasticity/plane_elastoplasticity.py:67-71``,
``tests/test_FeFp_jax.py:17-19``).  ``mpmath.findroot`` at 40 digits, tangents by central differences of the solution map
(step 1e-18 relative).  ``tests/test_oracle_jaxmat_form_mp.py`` bounds the distance between the canonical arithmetic
(what the CUDA kernels compute, bit for bit) and the exact solution of the equations jaxmat solves.  Vectors produced by
jaxmat itself are still what would turn "formulation-pinned" into "pinned"."""

import mpmath as mp

mp.mp.dps = 40
R2 = mp.sqrt(2)
IDX9 = ((0, 3, 5), (4, 1, 7), (6, 8, 2))  # reference 9-vector [11,22,33,12,21,13,31,23,32] (utils.py:173-186)


def fb(x, y):
    return x + y - mp.sqrt(x * x + y * y)


def sigma_y(p, props):
    sig0 = mp.mpf(props["sig0"])
    return (sig0 + mp.mpf(props.get("H", 0.0)) * p
            + (mp.mpf(props.get("sigu", props["sig0"])) - sig0) * (1 - mp.exp(-mp.mpf(props.get("b", 0.0)) * p)))


def _lame(props):
    E, nu = mp.mpf(props["E"]), mp.mpf(props["nu"])
    return E, E * nu / (1 + nu) / (1 - 2 * nu), E / 2 / (1 + nu), E / (3 * (1 - 2 * nu))


# ---- small strain ----------------------------------------------------------------------------------------------------
def j2_point(eps, eps_old, sig_old, p_old, props, dp0=0.0):
    """Returns ``dict(stress, p, depsp)`` (mpf lists) for Mandel 6-vectors ``eps, eps_old, sig_old``."""
    E, lam, mu, _ = _lame(props)
    de = [mp.mpf(a) - mp.mpf(b) for a, b in zip(eps, eps_old)]
    tr = de[0] + de[1] + de[2]
    sig_el = [mp.mpf(s) + 2 * mu * de[i] + (lam * tr if i < 3 else 0) for i, s in enumerate(sig_old)]
    pm = (sig_el[0] + sig_el[1] + sig_el[2]) / 3
    s = [sig_el[i] - (pm if i < 3 else 0) for i in range(6)]
    seq = mp.sqrt(mp.mpf(3) / 2 * sum(x * x for x in s))
    p_old = mp.mpf(p_old)
    f0 = seq - sigma_y(p_old, props)
    if f0 <= 0:
        dp = mp.mpf(0)  # FB(-f/E, 0) = 0 exactly for f <= 0
    else:
        dp = mp.findroot(lambda d: fb(-(seq - 3 * mu * d - sigma_y(p_old + d, props)) / E, d),
                         mp.mpf(dp0) if dp0 else f0 / (3 * mu), tol=mp.mpf(10) ** -60, verify=False)
    depsp = [mp.mpf(3) / 2 * dp * x / seq if seq != 0 else mp.mpf(0) for x in s]
    dee = [de[i] - depsp[i] for i in range(6)]
    tre = dee[0] + dee[1] + dee[2]
    sig = [mp.mpf(so) + 2 * mu * dee[i] + (lam * tre if i < 3 else 0) for i, so in enumerate(sig_old)]
    return dict(stress=sig, p=p_old + dp, depsp=depsp, dp=dp)


def j2_tangent(eps, eps_old, sig_old, p_old, props, dp0=0.0):
    scale = max(abs(mp.mpf(x)) for x in eps) or mp.mpf(1)
    h = scale * mp.mpf(10) ** -18
    Ct = mp.zeros(6, 6)
    for j in range(6):
        ep, em = [mp.mpf(x) for x in eps], [mp.mpf(x) for x in eps]
        ep[j] += h
        em[j] -= h
        sp = j2_point(ep, eps_old, sig_old, p_old, props, dp0)["stress"]
        sm = j2_point(em, eps_old, sig_old, p_old, props, dp0)["stress"]
        for i in range(6):
            Ct[i, j] = (sp[i] - sm[i]) / (2 * h)
    return Ct


# ---- finite strain ---------------------------------------------------------------------------------------------------
def _mat9(v):
    return mp.matrix([[mp.mpf(v[IDX9[i][j]]) for j in range(3)] for i in range(3)])


def _sym6(v):
    v = [mp.mpf(x) for x in v]
    return mp.matrix([[v[0], v[3] / R2, v[4] / R2], [v[3] / R2, v[1], v[5] / R2], [v[4] / R2, v[5] / R2, v[2]]])


def fefp_point(F9, F9_old, be_bar_old6, p_old, props, start=None):
    """Returns ``dict(PK1 (9-vector), p, be_bar (Mandel 6), plastic)``; ``start``: optional ``(dp, be_bar Mandel 6)`` guess."""
    E, lam, mu, kappa = _lame(props)
    I = mp.eye(3)
    F = _mat9(F9)
    f = F * mp.inverse(_mat9(F9_old))
    fbar = f * mp.det(f) ** (-mp.mpf(1) / 3)
    Btr = fbar * _sym6(be_bar_old6) * fbar.T
    p_old = mp.mpf(p_old)

    def unpack(x):
        return x[0], mp.matrix([[x[1], x[4], x[5]], [x[4], x[2], x[6]], [x[5], x[6], x[3]]])

    def residual(*x):
        dp, be = unpack(x)
        tr = be[0, 0] + be[1, 1] + be[2, 2]
        s = mu * (be - tr / 3 * I)
        seq = mp.sqrt(mp.mpf(3) / 2 * sum(s[i, j] ** 2 for i in range(3) for j in range(3)))
        fy = seq - sigma_y(p_old + dp, props)
        d = be - Btr
        trd = d[0, 0] + d[1, 1] + d[2, 2]
        R = d - trd / 3 * I + (mp.mpf(2) / 3 * dp * tr / seq) * mp.mpf(3) / 2 * s + I * (mp.det(be) - 1)
        return [fb(-fy / E, dp), R[0, 0], R[1, 1], R[2, 2], R[0, 1], R[0, 2], R[1, 2]]

    if start is not None:
        b = _sym6(start[1])
        x0 = [mp.mpf(start[0]), b[0, 0], b[1, 1], b[2, 2], b[0, 1], b[0, 2], b[1, 2]]
    else:
        x0 = [mp.mpf(0), Btr[0, 0], Btr[1, 1], Btr[2, 2], Btr[0, 1], Btr[0, 2], Btr[1, 2]]
    x = mp.findroot(residual, x0, tol=mp.mpf(10) ** -60, maxsteps=80, verify=False)
    res = residual(*x)
    assert max(abs(r) for r in res) < mp.mpf(10) ** -28, "40-digit solve did not converge"
    dp, be = unpack([x[i] for i in range(7)])
    J = mp.det(F)
    tr = be[0, 0] + be[1, 1] + be[2, 2]
    tau = mu * (be - tr / 3 * I) + kappa / 2 * (J * J - 1) * I
    P = tau * mp.inverse(F).T
    P9 = [None] * 9
    for i in range(3):
        for j in range(3):
            P9[IDX9[i][j]] = P[i, j]
    be6 = [be[0, 0], be[1, 1], be[2, 2], R2 * be[0, 1], R2 * be[0, 2], R2 * be[1, 2]]
    return dict(PK1=P9, p=p_old + dp, be_bar=be6, dp=dp)


def fefp_tangent(F9, F9_old, be_bar_old6, p_old, props, start=None):
    h = mp.mpf(10) ** -18
    Ct = mp.zeros(9, 9)
    for k in range(9):
        Fp, Fm = [mp.mpf(x) for x in F9], [mp.mpf(x) for x in F9]
        Fp[k] += h
        Fm[k] -= h
        Pp = fefp_point(Fp, F9_old, be_bar_old6, p_old, props, start)["PK1"]
        Pm = fefp_point(Fm, F9_old, be_bar_old6, p_old, props, start)["PK1"]
        for i in range(9):
            Ct[i, k] = (Pp[i] - Pm[i]) / (2 * h)
    return Ct
