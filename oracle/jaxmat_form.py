"""A second, independently written CPU restatement of the two jaxmat behaviours on the hot path -- in jaxmat's OWN
formulation.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The canonical oracle (``oracle/small_strain.py``, ``oracle/fefp.py``, mirrored by the CUDA kernels) branches on the trial
yield function, solves reduced scalar / 2x2 systems and writes the consistent tangent in closed form.  The reference
does none of that: ``JAXMaterial`` (``dolfinx_materials/jaxmat.py:147-164``) differentiates, with ``jax.jacfwd``, the
map ``gradient -> stress`` that jaxmat's behaviours define through a *branch-free* local problem

* ``vonMisesIsotropicHardening`` (call site ``demos/jax/elastoplasticity/plane_elastoplasticity.py:67-71``):
  unknown ``dp`` with ``FB(-f/E, dp) = 0``, ``f = seq_el - 3 mu dp - sigma_Y(p_old + dp)``,
  ``FB(x, y) = x + y - sqrt(x^2 + y^2)`` (Fischer-Burmeister complementarity: no elastic / plastic branch);
* ``FeFpJ2Plasticity`` (call sites ``tests/test_FeFp_jax.py:17-19``,
  ``demos/jax/finite_strain_elastoplasticity/finite_strain_elastoplasticity.py:165-181``): seven unknowns ``(dp, be_bar)``
  with ``FB(-f_y/E, dp) = 0`` and ``dev(be - be_tr) + 2/3 dp tr(be) n + 1 (det be - 1) = 0``;

with the local solve differentiated implicitly.  This module restates exactly that, sharing no code with the canonical
oracle: the branch-free residuals are written once for complex arguments, solved for every point (elastic ones too) by
a damped Newton whose Jacobian comes from the complex-step derivative of the residual, and the tangent is the
complex-step derivative of the converged stress map -- the *definition* of the reference's ``jacfwd`` result.
``tests/test_oracle_jaxmat_form.py`` holds the canonical oracle to it over the full golden histories: identical active
sets and rtol 1e-10 on stress, state and tangent.  That narrows "parity unpinned" to "formulation-pinned" -- only
vectors produced by jaxmat itself (``tests/golden/make_golden_jaxmat.py``) can close it.
"""

import numpy as np

H_CS = 1e-30  # complex-step size: derivative exact to rounding
R2 = np.sqrt(2.0)
ACTIVE_DP = 1e-13  # dp of an elastic point solves FB to ~1e-20 (never exactly 0): the active set is dp above round-off

# position of tensor component (i, j) in the reference's 9-vector [11,22,33,12,21,13,31,23,32] (utils.py:173-186)
IDX9 = ((0, 3, 5), (4, 1, 7), (6, 8, 2))


def fb(x, y):
    """Fischer-Burmeister function: fb(x, y) = 0  <=>  x >= 0, y >= 0, x y = 0."""
    return x + y - np.sqrt(x * x + y * y)


def sigma_y(p, props):
    """sig0 + H p + (sigu - sig0)(1 - exp(-b p)) as the demos write it (tests/test_FeFp_jax.py:14-15)."""
    sig0 = props["sig0"]
    return sig0 + props.get("H", 0.0) * p + (props.get("sigu", sig0) - sig0) * (1.0 - np.exp(-props.get("b", 0.0) * p))


def _newton(res, x0, tol=1e-14, max_it=60):
    """Damped Newton on a batch of small nonlinear systems.  ``res(x)``: (n, m) complex -> (n, m) complex; the Jacobian
    is the complex-step derivative of ``res`` (exact).  Returns the real solution (n, m)."""
    x = np.array(x0, dtype=float)
    n, m = x.shape
    r = res(x.astype(complex)).real
    nr = np.abs(r).max(axis=1)
    for _ in range(max_it):
        live = nr > tol
        if not live.any():
            break
        J = np.zeros((n, m, m))
        for k in range(m):
            xp = x.astype(complex)
            xp[:, k] += 1j * H_CS
            J[:, :, k] = res(xp).imag / H_CS
        dx = np.zeros_like(x)
        dx[live] = -np.linalg.solve(J[live], r[live][:, :, None])[:, :, 0]
        t = np.ones(n)
        for _ in range(30):  # simple decrease
            xn = x + t[:, None] * dx
            rn = res(xn.astype(complex)).real
            nn = np.abs(rn).max(axis=1)
            bad = live & ~(nn < nr)
            if not bad.any():
                break
            t = np.where(bad, 0.5 * t, t)
        x, r, nr = xn, rn, nn
    return x


def _complex_root(res, x_real, sweeps=4):
    """Root of the complex-analytic residual next to the real root, for a complex-perturbed parameter: fixed point with
    the real Jacobian at the real root -- exact to first order in the imaginary part after a few sweeps (the implicit
    differentiation the reference gets from optimistix)."""
    n, m = x_real.shape
    J = np.zeros((n, m, m))
    for k in range(m):
        xp = x_real.astype(complex)
        xp[:, k] += 1j * H_CS
        J[:, :, k] = res(xp, real_params=True).imag / H_CS
    Ji = np.linalg.inv(J)
    x = x_real.astype(complex)
    for _ in range(sweeps):
        x = x - np.einsum("nij,nj->ni", Ji, res(x))
    return x


# ---- small strain: vonMisesIsotropicHardening ------------------------------------------------------------------------
def _j2_stress(eps, st, props, dp_real=None):
    """(sigma, dp, epsp, seq_el) for (possibly complex) strains ``eps`` (n, 6); dp from FB(-f/E, dp) = 0."""
    E, nu = props["E"], props["nu"]
    lam, mu = E * nu / (1 + nu) / (1 - 2 * nu), E / 2 / (1 + nu)
    C = 2 * mu * np.eye(6)
    C[:3, :3] += lam
    sig_el = st["stress"] + (eps - st["strain"]) @ C.T
    s = sig_el.copy()
    s[:, :3] -= sig_el[:, :3].sum(axis=1, keepdims=True) / 3
    seq_el = np.sqrt(1.5 * (s * s).sum(axis=1))
    seq_el = np.where(np.abs(seq_el) > 1e-8, seq_el, 1e-8)  # jaxmat clips the norm away from 0
    p_old = st["p"]

    def res(x, real_params=False):
        sq = seq_el.real if real_params else seq_el
        dp = x[:, 0]
        f = sq - 3 * mu * dp - sigma_y(p_old + dp, props)
        return fb(-f / E, dp)[:, None]

    if dp_real is None:
        dp_real = _newton(res, np.zeros((len(eps), 1)))
    dp = _complex_root(res, dp_real)[:, 0] if np.iscomplexobj(eps) else dp_real[:, 0]
    depsp = 1.5 * dp[:, None] * s / seq_el[:, None]
    sig = st["stress"] + (eps - st["strain"] - depsp) @ C.T
    return sig, dp, depsp, dp_real


def j2_integrate(eps, state, props):
    """jaxmat-form update of a batch: returns ``stress, p, epsp, Ct, flag`` (``Ct = d stress / d strain`` by the
    complex-step derivative of the converged map, i.e. what ``jacfwd`` + implicit differentiation give)."""
    eps = np.asarray(eps, dtype=float)
    n = len(eps)
    st = {"strain": np.asarray(state["strain"], float).reshape(n, 6), "stress": np.asarray(state["stress"], float).reshape(n, 6),
          "p": np.asarray(state["p"], float).reshape(n), "epsp": np.asarray(state["epsp"], float).reshape(n, 6)}
    sig, dp, depsp, dp_real = _j2_stress(eps, st, props)
    Ct = np.zeros((n, 6, 6))
    for k in range(6):
        e = eps.astype(complex)
        e[:, k] += 1j * H_CS
        Ct[:, :, k] = _j2_stress(e, st, props, dp_real)[0].imag / H_CS
    return {"stress": sig, "p": st["p"] + dp, "epsp": st["epsp"] + depsp, "Ct": Ct, "flag": (dp > ACTIVE_DP).astype(np.uint8)}


# ---- finite strain: FeFpJ2Plasticity ---------------------------------------------------------------------------------
def _mat9(v):
    return np.stack([np.stack([v[:, IDX9[i][j]] for j in range(3)], axis=1) for i in range(3)], axis=1)


def _vec9(a):
    out = np.zeros((a.shape[0], 9), dtype=a.dtype)
    for i in range(3):
        for j in range(3):
            out[:, IDX9[i][j]] = a[:, i, j]
    return out


def _sym6_to_mat(v):
    return np.stack([np.stack([v[:, 0], v[:, 3] / R2, v[:, 4] / R2], axis=1),
                     np.stack([v[:, 3] / R2, v[:, 1], v[:, 5] / R2], axis=1),
                     np.stack([v[:, 4] / R2, v[:, 5] / R2, v[:, 2]], axis=1)], axis=1)


def _mat_to_sym6(a):
    return np.stack([a[:, 0, 0], a[:, 1, 1], a[:, 2, 2], R2 * a[:, 0, 1], R2 * a[:, 0, 2], R2 * a[:, 1, 2]], axis=1)


def _det(a):
    return (a[:, 0, 0] * (a[:, 1, 1] * a[:, 2, 2] - a[:, 1, 2] * a[:, 2, 1])
            - a[:, 0, 1] * (a[:, 1, 0] * a[:, 2, 2] - a[:, 1, 2] * a[:, 2, 0])
            + a[:, 0, 2] * (a[:, 1, 0] * a[:, 2, 1] - a[:, 1, 1] * a[:, 2, 0]))


def _inv(a):
    """3x3 inverse by cofactors (complex-analytic, unlike a pivoting solver's branch choices)."""
    c = np.empty_like(a)
    for i in range(3):
        for j in range(3):
            i1, i2, j1, j2 = (i + 1) % 3, (i + 2) % 3, (j + 1) % 3, (j + 2) % 3
            c[:, j, i] = a[:, i1, j1] * a[:, i2, j2] - a[:, i1, j2] * a[:, i2, j1]
    return c / _det(a)[:, None, None]


def _unpack7(x):
    be = np.stack([np.stack([x[:, 1], x[:, 4], x[:, 5]], axis=1), np.stack([x[:, 4], x[:, 2], x[:, 6]], axis=1),
                   np.stack([x[:, 5], x[:, 6], x[:, 3]], axis=1)], axis=1)
    return x[:, 0], be


def _fefp_pk1(F9, st, props, x_real=None):
    E, nu = props["E"], props["nu"]
    mu, kappa = E / 2 / (1 + nu), E / (3 * (1 - 2 * nu))
    I = np.eye(3)[None]
    F = _mat9(F9)
    f = F @ _inv(_mat9(st["F"]).astype(F.dtype))
    fbar = f * (_det(f) ** (-1.0 / 3.0))[:, None, None]
    Btr = fbar @ _sym6_to_mat(st["be_bar"]).astype(F.dtype) @ np.swapaxes(fbar, 1, 2)
    p_old = st["p"]

    def res(x, real_params=False):
        B = Btr.real if real_params else Btr
        dp, be = _unpack7(x)
        tr = be[:, 0, 0] + be[:, 1, 1] + be[:, 2, 2]
        s = mu * (be - tr[:, None, None] / 3 * I)
        seq = np.sqrt(1.5 * (s * s).sum(axis=(1, 2)))
        seq = np.where(np.abs(seq) > 1e-8, seq, 1e-8)
        fy = seq - sigma_y(p_old + dp, props)
        d = be - B
        trd = d[:, 0, 0] + d[:, 1, 1] + d[:, 2, 2]
        R = d - trd[:, None, None] / 3 * I + (2.0 / 3.0 * dp * tr / seq)[:, None, None] * 1.5 * s + I * (_det(be) - 1)[:, None, None]
        return np.stack([fb(-fy / E, dp), R[:, 0, 0], R[:, 1, 1], R[:, 2, 2], R[:, 0, 1], R[:, 0, 2], R[:, 1, 2]], axis=1)

    if x_real is None:
        B0 = Btr.real
        x0 = np.stack([np.zeros(len(F9)), B0[:, 0, 0], B0[:, 1, 1], B0[:, 2, 2], B0[:, 0, 1], B0[:, 0, 2], B0[:, 1, 2]], axis=1)
        x_real = _newton(res, x0)
    x = _complex_root(res, x_real) if np.iscomplexobj(F9) else x_real
    dp, be = _unpack7(x)
    J = _det(F)
    tr = be[:, 0, 0] + be[:, 1, 1] + be[:, 2, 2]
    tau = mu * (be - tr[:, None, None] / 3 * I) + (kappa / 2 * (J * J - 1))[:, None, None] * I
    P = tau @ np.swapaxes(_inv(F), 1, 2)
    return _vec9(P), dp, be, x_real


def fefp_integrate(F, state, props):
    """jaxmat-form FeFp update of a batch: ``PK1, p, be_bar, Ct = dPK1/dF, flag``."""
    F = np.asarray(F, dtype=float)
    n = len(F)
    st = {"F": np.asarray(state["F"], float).reshape(n, 9), "p": np.asarray(state["p"], float).reshape(n),
          "be_bar": np.asarray(state["be_bar"], float).reshape(n, 6)}
    P, dp, be, x_real = _fefp_pk1(F, st, props)
    Ct = np.zeros((n, 9, 9))
    for k in range(9):
        Fc = F.astype(complex)
        Fc[:, k] += 1j * H_CS
        Ct[:, :, k] = _fefp_pk1(Fc, st, props, x_real)[0].imag / H_CS
    return {"PK1": P, "p": st["p"] + dp, "be_bar": _mat_to_sym6(be), "Ct": Ct, "flag": (dp > ACTIVE_DP).astype(np.uint8)}
