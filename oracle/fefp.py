"""numpy fp64 restatement of the finite-strain hot path: multiplicative (Fe.Fp) J2 plasticity with the
isochoric elastic left Cauchy-Green tensor ``be_bar`` and the cumulated plastic strain ``p`` as internal
state.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  **Parity unpinned**: the behaviour
(``jaxmat.materials.FeFpJ2Plasticity``) lives in the un-vendored jaxmat package and the reference's only
test of it (``tests/test_FeFp_jax.py:6-33``) asserts nothing; what follows restates the published
algorithm and is checked by self-consistency in ``tests/test_oracle_fefp.py``.

Protocol (reference ``dolfinx_materials/jaxmat.py:166-193``): gradient ``F`` (9), flux ``PK1`` (9), internal
state ``p`` (1), ``be_bar`` (6, Mandel, identity-initialised: ``finite_strain_elastoplasticity.py:181``);
non-symmetric tensors are ordered ``[11,22,33,12,21,13,31,23,32]`` (``dolfinx_materials/utils.py:173-186``);
``Ct[a, b] = dPK1_a / dF_b`` (``quadrature_map.py:94-104``).

Algorithm
---------
``f = F F_old^-1``, ``B = det(f)^(-2/3) f be_old f^T`` (trial), ``D = dev B``, ``seq_tr = mu sqrt(3/2 D:D)``.
Active set: ``seq_tr - sigma_Y(p_old) > 0``.  The 7-unknown system of the reference formulation
(``dev(be - B) + 2/3 dp tr(be) n + 1 (det be - 1) = 0`` with the yield condition) has ``dev be`` parallel to
``D``, so with ``be = alpha D + t 1`` and ``alpha = 1 - 3 mu t dp / seq_tr`` it reduces exactly to two scalars:

    r1(dp, t) = seq_tr - 3 mu t dp - sigma_Y(p_old + dp) = 0
    r2(dp, t) = t^3 - 1/2 alpha^2 (D:D) t + alpha^3 det(D) - 1 = 0          (det be = 1)

solved by a 2x2 Newton from ``(0, tr B / 3)``.  ``tau = mu alpha D + kappa/2 (J^2 - 1) 1``, ``PK1 = tau F^-T``.
The tangent is the exact linearisation (implicit differentiation of the local solve), assembled column by
column in closed form.  Elastic points keep ``be = B``.
Every expression is written component-wise in a fixed order (``oracle/canon.py``); the CUDA kernel
``dxm_fefp_kernel`` follows the same order.
"""

import numpy as np

from .canon import dot3 as _dot3
from .canon import exp_c, fma, fms, fnma

NEWTON_CAP = 25
NEWTON_RTOL = 1e-12
RSQRT2 = 0.70710678118654752440
SQRT2 = 1.41421356237309504880

# position of tensor component (i, j) in the reference's 9-vector (utils.py:173-186)
IDX9 = ((0, 3, 5), (4, 1, 7), (6, 8, 2))


def _col(a, n):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 0:
        return np.full(n, float(a))
    return a.reshape(n)


THIRD = 1.0 / 3.0


def rcbrt_c(x):
    """x^(-1/3) for x > 0 from exactly rounded operations only, division free: frexp range reduction
    to [0.5, 4), linear initial guess, 6 Newton steps y <- y (4 - x y^3) / 3, exact rescaling
    (NaN for x <= 0).  Accurate to 2 ulp."""
    x = np.asarray(x, dtype=np.float64)
    ok = x > 0.0
    xs = np.where(ok, x, 1.0)
    m, e = np.frexp(xs)
    q = np.floor_divide(e, 3)
    r = e - 3 * q
    xr = np.ldexp(m, r)
    y = fnma(0.15, xr, 1.2)
    for _ in range(6):
        y = (y * fnma(xr, (y * y) * y, 4.0)) * THIRD
    y = np.ldexp(y, -q)
    return np.where(ok, y, np.nan)


def _unpack9(v):
    return [[v[:, IDX9[i][j]] for j in range(3)] for i in range(3)]


def _det3(A):
    m0 = fms(A[1][1], A[2][2], A[1][2] * A[2][1])
    m1 = fms(A[1][0], A[2][2], A[1][2] * A[2][0])
    m2 = fms(A[1][0], A[2][1], A[1][1] * A[2][0])
    return fma(A[0][2], m2, fnma(A[0][1], m1, A[0][0] * m0))


def _inv3(A):
    c = [[None] * 3 for _ in range(3)]
    c[0][0] = fms(A[1][1], A[2][2], A[1][2] * A[2][1])
    c[0][1] = fms(A[0][2], A[2][1], A[0][1] * A[2][2])
    c[0][2] = fms(A[0][1], A[1][2], A[0][2] * A[1][1])
    c[1][0] = fms(A[1][2], A[2][0], A[1][0] * A[2][2])
    c[1][1] = fms(A[0][0], A[2][2], A[0][2] * A[2][0])
    c[1][2] = fms(A[0][2], A[1][0], A[0][0] * A[1][2])
    c[2][0] = fms(A[1][0], A[2][1], A[1][1] * A[2][0])
    c[2][1] = fms(A[0][1], A[2][0], A[0][0] * A[2][1])
    c[2][2] = fms(A[0][0], A[1][1], A[0][1] * A[1][0])
    det = fma(A[0][2], c[2][0], fma(A[0][1], c[1][0], A[0][0] * c[0][0]))
    rdet = 1.0 / det
    return [[c[i][j] * rdet for j in range(3)] for i in range(3)], det


def integrate(F, state, props, newton_cap=NEWTON_CAP, rtol=NEWTON_RTOL):
    """One batched update.  ``F`` (n,9); ``state`` dict ``F`` (n,9), ``PK1`` (n,9) (unused), ``p`` (n,),
    ``be_bar`` (n,6); ``props``: ``E, nu, sig0`` and optionally ``H, sigu, b`` (scalars or (n,) arrays).
    Returns ``F, PK1, p, be_bar`` (= s1), ``Ct`` (n,9,9), ``flag``, ``n_iter``, ``resid``, ``fail``."""
    F = np.ascontiguousarray(F, dtype=np.float64)
    n = F.shape[0]
    Fo = np.asarray(state["F"], dtype=np.float64).reshape(n, 9)
    beo = np.asarray(state["be_bar"], dtype=np.float64).reshape(n, 6)
    p_old = np.asarray(state["p"], dtype=np.float64).reshape(n)

    E = _col(props["E"], n)
    nu = _col(props["nu"], n)
    sig0 = _col(props["sig0"], n)
    H = _col(props.get("H", 0.0), n)
    sigu = _col(props.get("sigu", props["sig0"]), n)
    b = _col(props.get("b", 0.0), n)

    with np.errstate(all="ignore"):
        mu = E / 2 / (1 + nu)
        kappa = E / (3 * (1 - 2 * nu))
        threemu = 3.0 * mu
        dsu = sigu - sig0
        dsu = np.where(np.isfinite(dsu), dsu, 0.0)
        bdsu = b * dsu

        A = _unpack9(F)
        Ao = _unpack9(Fo)
        Bo = [[None] * 3 for _ in range(3)]
        Bo[0][0], Bo[1][1], Bo[2][2] = beo[:, 0], beo[:, 1], beo[:, 2]
        Bo[0][1] = Bo[1][0] = beo[:, 3] * RSQRT2
        Bo[0][2] = Bo[2][0] = beo[:, 4] * RSQRT2
        Bo[1][2] = Bo[2][1] = beo[:, 5] * RSQRT2

        # a singular / inverted elastic state (det be_bar <= 0, e.g. a be_bar left at zero instead of the identity)
        # would give PK1 = 0 with every other check green: counted as a failed point
        det_bo = _det3(Bo)

        # ---- trial state -------------------------------------------------------------------
        Aoi, _ = _inv3(Ao)
        f = [[_dot3(A[i][0], Aoi[0][j], A[i][1], Aoi[1][j], A[i][2], Aoi[2][j]) for j in range(3)] for i in range(3)]
        Jf = _det3(f)
        rc = rcbrt_c(Jf)
        s23 = rc * rc
        M = [[_dot3(f[i][0], Bo[0][j], f[i][1], Bo[1][j], f[i][2], Bo[2][j]) for j in range(3)] for i in range(3)]
        B = [[None] * 3 for _ in range(3)]
        for i in range(3):
            for j in range(i, 3):
                B[i][j] = s23 * _dot3(M[i][0], f[j][0], M[i][1], f[j][1], M[i][2], f[j][2])
                B[j][i] = B[i][j]
        t0 = ((B[0][0] + B[1][1]) + B[2][2]) * THIRD
        D = [[B[i][j] - t0 if i == j else B[i][j] for j in range(3)] for i in range(3)]
        dd = fma(2.0, _dot3(D[0][1], D[0][1], D[0][2], D[0][2], D[1][2], D[1][2]),
                 _dot3(D[0][0], D[0][0], D[1][1], D[1][1], D[2][2], D[2][2]))
        d3 = _det3(D)
        seq = mu * np.sqrt(1.5 * dd)
        rseq = 1.0 / seq

        e0 = exp_c(-(b * p_old))
        sy0 = fma(dsu, 1.0 - e0, fma(H, p_old, sig0))
        ftr = seq - sy0
        flag = ftr > 0.0

        # ---- local 2x2 Newton --------------------------------------------------------------------
        c = threemu * rseq
        dp = np.zeros(n)
        t = t0.copy()
        ecur = e0.copy()
        n_iter = np.zeros(n, dtype=np.int32)
        resid = np.zeros(n)
        fail = np.zeros(n, dtype=bool)
        active = flag.copy()
        tol1 = rtol * seq
        for it in range(newton_cap + 1):
            if not active.any():
                break
            ct = c * t
            tmt = threemu * t
            alpha = fnma(ct, dp, 1.0)
            p = p_old + dp
            sy = fma(dsu, 1.0 - ecur, fma(H, p, sig0))
            r1 = fnma(tmt, dp, seq) - sy
            a2 = alpha * alpha
            ha2 = 0.5 * a2
            tt = t * t
            r2 = fnma(ha2, dd * t, tt * t) + fms(a2 * alpha, d3, 1.0)
            conv = (np.abs(r1) <= tol1) & (np.abs(r2) <= rtol)
            resid = np.where(active & conv, np.abs(r1), resid)
            active = active & ~conv
            if it == newton_cap:
                fail |= active
                resid = np.where(active, np.abs(r1), resid)
                break
            dsy = fma(bdsu, ecur, H)
            g = fnma(alpha * dd, t, 3.0 * (a2 * d3))
            J11 = -tmt - dsy
            J12 = -(threemu * dp)
            cdp = c * dp
            J21 = -(g * ct)
            J22 = fnma(g, cdp, fnma(ha2, dd, 3.0 * tt))
            rdet = 1.0 / fms(J11, J22, J12 * J21)
            dp_new = fma(fms(J12, r2, r1 * J22), rdet, dp)
            t_new = fma(fms(J21, r1, J11 * r2), rdet, t)
            dp = np.where(active, dp_new, dp)
            t = np.where(active, t_new, t)
            e_new = exp_c(-(b * (p_old + dp)))
            ecur = np.where(active, e_new, ecur)
            n_iter = n_iter + active.astype(np.int32)

        dp = np.where(flag, dp, 0.0)
        t = np.where(flag, t, t0)
        alpha = np.where(flag, fnma(c * t, dp, 1.0), 1.0)
        p_new = p_old + dp

        # ---- new state -----------------------------------------------------------------------------
        be = [None] * 6
        for i in range(3):
            be[i] = np.where(flag, fma(alpha, D[i][i], t), B[i][i])
        be[3] = (alpha * D[0][1]) * SQRT2
        be[4] = (alpha * D[0][2]) * SQRT2
        be[5] = (alpha * D[1][2]) * SQRT2

        # ---- stress: tau = mu alpha D + pvol 1, PK1 = tau F^-T = mu alpha (D F^-T) + pvol F^-T ----------
        Ai, Jd = _inv3(A)
        muA = mu * alpha
        pvol = (0.5 * kappa) * fms(Jd, Jd, 1.0)
        DA = [[_dot3(D[i][0], Ai[j][0], D[i][1], Ai[j][1], D[i][2], Ai[j][2]) for j in range(3)] for i in range(3)]
        P = [[fma(muA, DA[i][j], pvol * Ai[j][i]) for j in range(3)] for i in range(3)]

        # ---- local-solve sensitivities: d(alpha) = al1 * (D:dD) + al2 * (D^2:dD) --------------------------
        sq1 = (1.5 * (mu * mu)) * rseq
        a2 = alpha * alpha
        dsy = fma(bdsu, ecur, H)
        g = fnma(alpha * dd, t, 3.0 * (a2 * d3))
        ct = c * t
        cdp = c * dp
        J11 = -(threemu * t) - dsy
        J12 = -(threemu * dp)
        J21 = -(g * ct)
        J22 = fnma(g, cdp, fnma(0.5 * a2, dd, 3.0 * (t * t)))
        rdet = 1.0 / fms(J11, J22, J12 * J21)
        oma = (1.0 - alpha) * rseq
        b21 = fms(g * oma, sq1, a2 * t)
        b22 = a2 * alpha
        p1 = -(fms(sq1, J22, J12 * b21) * rdet)
        t1 = -(fms(J11, b21, J21 * sq1) * rdet)
        p2 = (J12 * b22) * rdet
        t2 = -((J11 * b22) * rdet)
        al1 = np.where(flag, fnma(cdp, t1, fnma(ct, p1, oma * sq1)), 0.0)
        al2 = np.where(flag, fnma(cdp, t2, -(ct * p2)), 0.0)

        # ---- tangent, column (k, l) = d/dF_kl:  with w = row l of F^-1, v = B w = D w + t0 w -------------
        #   dP_ij = cD (D F^-T)_ij + cI F^-T_ij + delta_ik mu alpha (F^-1 v)_j + hs w_i F^-1_jk
        kJ2 = kappa * (Jd * Jd)
        c23dd = (2.0 / 3.0) * dd
        twod3 = 2.0 * d3
        c23muA = (2.0 / 3.0) * muA
        hs = fms(muA, t0, pvol)
        Ct = np.zeros((n, 9, 9))
        for l in range(3):
            w = [Ai[l][0], Ai[l][1], Ai[l][2]]
            v = [fma(t0, w[i], _dot3(D[i][0], w[0], D[i][1], w[1], D[i][2], w[2])) for i in range(3)]
            u = [_dot3(D[i][0], v[0], D[i][1], v[1], D[i][2], v[2]) for i in range(3)]
            z = [_dot3(D[i][0], u[0], D[i][1], u[1], D[i][2], u[2]) for i in range(3)]
            my = [muA * _dot3(Ai[j][0], v[0], Ai[j][1], v[1], Ai[j][2], v[2]) for j in range(3)]
            hw = [hs * w[i] for i in range(3)]
            for k in range(3):
                a1 = fnma(c23dd, w[k], 2.0 * u[k])
                a2p = fnma(twod3, w[k], fnma(c23dd, v[k], 2.0 * z[k]))
                cD = fnma(c23muA, w[k], np.where(flag, mu * fma(al2, a2p, al1 * a1), 0.0))
                cI = fnma(c23muA, v[k], kJ2 * w[k])
                col = IDX9[k][l]
                for i in range(3):
                    for j in range(3):
                        val = fma(hw[i], Ai[j][k], fma(cI, Ai[j][i], cD * DA[i][j]))
                        if i == k:
                            val = val + my[j]
                        Ct[:, IDX9[i][j], col] = val

        chk = (seq + np.abs(Jd)) + p_new
        for i in range(6):
            chk = chk + np.abs(be[i])
        for i in range(3):
            for j in range(3):
                chk = chk + np.abs(P[i][j])
        fail |= ~np.isfinite(chk) | ~(det_bo > 0.0)

    PK1 = np.empty((n, 9))
    for i in range(3):
        for j in range(3):
            PK1[:, IDX9[i][j]] = P[i][j]
    return {
        "F": F,
        "PK1": PK1,
        "p": p_new,
        "be_bar": np.stack(be, axis=1),
        "Ct": Ct,
        "flag": flag.astype(np.uint8),
        "n_iter": n_iter,
        "resid": resid,
        "fail": fail.astype(np.uint8),
    }


def virgin_state(n):
    """F = I, PK1 = 0, p = 0, be_bar = I (jaxmat ``init_state``; demo ``finite_strain_elastoplasticity.py:181``)."""
    F = np.zeros((n, 9))
    F[:, :3] = 1.0
    be = np.zeros((n, 6))
    be[:, :3] = 1.0
    return {"F": F, "PK1": np.zeros((n, 9)), "p": np.zeros(n), "be_bar": be}


def advance(out):
    return {k: out[k] for k in ("F", "PK1", "p", "be_bar")}
