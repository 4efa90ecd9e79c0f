"""The reference obtains the tangent as ``jax.jacfwd`` of the converged stress map, with the local solve
differentiated implicitly (``dolfinx_materials/jaxmat.py:147-155``) -- i.e. the EXACT derivative of the solution of
jaxmat's residual system.  jaxmat is not installable here (parity unpinned, see oracle/__init__.py), so these tests
re-create that definition independently of the oracle's closed-form tangents: the residual systems of SURVEY.md
A.3 / A.4 in their Fischer-Burmeister form are written directly in complex arithmetic, solved from the converged real
solution with a frozen real Jacobian, and differentiated by the complex-step method (no subtractive cancellation:
the derivative is exact to rounding).  The oracle's stress and tangent must agree with that at the north star's
rtol 1e-10 level."""
import numpy as np

from oracle import fefp, synth
from oracle import small_strain as ss

H_CS = 1e-30
R2 = np.sqrt(2.0)


def fb(x, y):
    return x + y - np.sqrt(x * x + y * y)


def mandel_to_tensor(v):
    return np.array([[v[0], v[3] / R2, v[4] / R2], [v[3] / R2, v[1], v[5] / R2], [v[4] / R2, v[5] / R2, v[2]]])


def tensor_to_mandel(a):
    return np.array([a[0, 0], a[1, 1], a[2, 2], R2 * a[0, 1], R2 * a[0, 2], R2 * a[1, 2]])


def complex_root(res, x_real, nfix=4):
    """root of the complex-analytic residual next to the real root: fixed-point with the frozen real Jacobian
    (finite differences), exact to first order in the imaginary perturbation after a few sweeps"""
    n = len(x_real)
    J = np.zeros((n, n))
    r0 = res(x_real.astype(complex)).real
    for k in range(n):
        h = 1e-7 * max(1.0, abs(x_real[k]))
        xp = x_real.astype(complex)
        xp[k] += h
        J[:, k] = (res(xp).real - r0) / h
    Ji = np.linalg.inv(J)
    x = x_real.astype(complex)
    for _ in range(nfix):
        x = x - Ji @ res(x)
    return x


# ---- small strain J2 + Voce, jaxmat formulation -------------------------------------------------------------
VOCE = dict(E=70e3, nu=0.3, sig0=350.0, sigu=500.0, b=1e3)


def voce_stress(eps, st_i, dp_guess):
    """sigma(eps) with dp from FB(-f/E, dp) = 0, all in complex arithmetic; returns (sigma Mandel, dp)"""
    E, nu = VOCE["E"], VOCE["nu"]
    lam, mu = E * nu / (1 + nu) / (1 - 2 * nu), E / 2 / (1 + nu)
    C = 2 * mu * np.eye(6)
    C[:3, :3] += lam
    sig_el = st_i["stress"] + C @ (eps - st_i["strain"])
    s = sig_el.copy()
    s[:3] -= sig_el[:3].sum() / 3
    seq_el = np.sqrt(1.5 * (s * s).sum())

    def res(x):
        dp = x[0]
        f = seq_el - 3 * mu * dp - (VOCE["sig0"] + (VOCE["sigu"] - VOCE["sig0"]) * (1 - np.exp(-VOCE["b"] * (st_i["p"] + dp))))
        return np.array([fb(-f / E, dp)])

    dp = complex_root(res, np.array([dp_guess]))[0]
    depsp = 1.5 * dp * s / seq_el
    return st_i["stress"] + C @ (eps - st_i["strain"] - depsp), dp


def test_voce_tangent_is_the_exact_derivative_of_the_fischer_burmeister_solution():
    n, K = 60, 4
    st = ss.zero_state(n)
    for k in range(1, K):
        st = ss.advance(ss.integrate(synth.strain(n, 0, 1.25e-2, k, K), st, VOCE))
    eps = synth.strain(n, 0, 1.25e-2, K, K)
    out = ss.integrate(eps, st, VOCE)
    idx = np.flatnonzero(out["flag"])[:20]
    assert len(idx) == 20
    worst_s = worst_c = 0.0
    for i in idx:
        st_i = {k: (v[i] if v.ndim > 1 else v[i]) for k, v in st.items()}
        dp0 = out["p"][i] - st["p"][i]
        sig, dp = voce_stress(eps[i].astype(complex), st_i, dp0)
        assert abs(dp.real - dp0) <= 1e-11 * max(dp0, 1e-6)  # the oracle Newton stops at |r| <= 1e-12 seq
        worst_s = max(worst_s, np.abs(sig.real - out["stress"][i]).max() / np.abs(out["stress"][i]).max())
        Ct = np.zeros((6, 6))
        for k in range(6):
            e = eps[i].astype(complex)
            e[k] += 1j * H_CS
            Ct[:, k] = voce_stress(e, st_i, dp0)[0].imag / H_CS
        worst_c = max(worst_c, np.abs(Ct - out["Ct"][i]).max() / np.abs(out["Ct"][i]).max())
    assert worst_s < 1e-11 and worst_c < 1e-10, (worst_s, worst_c)


# ---- finite strain FeFp, jaxmat formulation: unknowns (dp, be_bar) ------------------------------------------
FEFP = dict(E=70e3, nu=0.3, sig0=500.0, sigu=750.0, b=1000.0)
MU, KAPPA = 70e3 / 2 / 1.3, 70e3 / (3 * 0.4)


def mat9(v):
    return np.array([[v[fefp.IDX9[i][j]] for j in range(3)] for i in range(3)])


def vec9(a):
    out = np.zeros(9, dtype=a.dtype)
    for i in range(3):
        for j in range(3):
            out[fefp.IDX9[i][j]] = a[i, j]
    return out


def det3(a):
    return (a[0, 0] * (a[1, 1] * a[2, 2] - a[1, 2] * a[2, 1]) - a[0, 1] * (a[1, 0] * a[2, 2] - a[1, 2] * a[2, 0])
            + a[0, 2] * (a[1, 0] * a[2, 1] - a[1, 1] * a[2, 0]))


def fefp_pk1(F9, st_i, x_guess):
    F = mat9(F9)
    f = F @ np.linalg.inv(mat9(st_i["F"]))
    fbar = f * det3(f) ** (-1.0 / 3.0)
    Btr = fbar @ mandel_to_tensor(st_i["be_bar"]) @ fbar.T
    p_old = st_i["p"]
    I = np.eye(3)

    def unpack(x):
        return x[0], np.array([[x[1], x[4], x[5]], [x[4], x[2], x[6]], [x[5], x[6], x[3]]])

    def res(x):
        dp, be = unpack(x)
        s = MU * (be - np.trace(be) / 3 * I)
        seq = np.sqrt(1.5 * (s * s).sum())
        fy = seq - (FEFP["sig0"] + (FEFP["sigu"] - FEFP["sig0"]) * (1 - np.exp(-FEFP["b"] * (p_old + dp))))
        R = (be - Btr) - np.trace(be - Btr) / 3 * I + 2.0 / 3.0 * dp * np.trace(be) * 1.5 * s / seq + I * (det3(be) - 1)
        return np.array([fb(-fy / FEFP["E"], dp), R[0, 0], R[1, 1], R[2, 2], R[0, 1], R[0, 2], R[1, 2]])

    x = complex_root(res, x_guess)
    dp, be = unpack(x)
    J = det3(F)
    tau = MU * (be - np.trace(be) / 3 * I) + KAPPA / 2 * (J * J - 1) * I
    return vec9(tau @ np.linalg.inv(F).T), x


def test_fefp_tangent_is_the_exact_derivative_of_the_seven_unknown_solution():
    n, K = 200, 4
    st = fefp.virgin_state(n)
    for k in range(1, K):
        st = fefp.advance(fefp.integrate(synth.defgrad(n, 0, 6e-2, k, K), st, FEFP))
    F = synth.defgrad(n, 0, 6e-2, K, K)
    out = fefp.integrate(F, st, FEFP)
    idx = np.flatnonzero(out["flag"])[:12]
    assert len(idx) == 12
    worst_s = worst_c = 0.0
    for i in idx:
        st_i = {k: v[i] for k, v in st.items()}
        be = mandel_to_tensor(out["be_bar"][i])
        xg = np.array([out["p"][i] - st["p"][i], be[0, 0], be[1, 1], be[2, 2], be[0, 1], be[0, 2], be[1, 2]])
        P, x = fefp_pk1(F[i].astype(complex), st_i, xg)
        assert np.abs(x.real - xg).max() < 1e-11  # the oracle's reduced 2x2 solve is a root of the 7-unknown system
        worst_s = max(worst_s, np.abs(P.real - out["PK1"][i]).max() / np.abs(out["PK1"][i]).max())
        Ct = np.zeros((9, 9))
        for k in range(9):
            Fc = F[i].astype(complex)
            Fc[k] += 1j * H_CS
            Ct[:, k] = fefp_pk1(Fc, st_i, xg)[0].imag / H_CS
        worst_c = max(worst_c, np.abs(Ct - out["Ct"][i]).max() / np.abs(out["Ct"][i]).max())
    assert worst_s < 1e-11 and worst_c < 1e-10, (worst_s, worst_c)
