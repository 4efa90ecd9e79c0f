"""numpy fp64 restatement of the small-strain hot path: linear elasticity and J2 plasticity with
linear, Voce or mixed isotropic hardening.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

What it restates
----------------
* protocol: ``Material.integrate(gradients, dt)`` -> ``(flux, isv, Ct)`` with state carried from
  ``s0`` (reference ``dolfinx_materials/generic.py:176-189``); state fields and their order are the
  ones the jaxmat back-end exposes (``dolfinx_materials/jaxmat.py:166-193``): gradient ``strain``
  (6, Mandel), flux ``stress`` (6), internal state ``p`` (1), ``epsp`` (6).
* elasticity: ``C = lambda 1x1 + 2 mu I6`` (``python_materials/elasticity.py:12-24``).
* J2 radial return + consistent tangent: ``tests/mfront/IsotropicLinearHardeningPlasticity.mfront:49-77``,
  written in the stress-increment form jaxmat uses (``sigma_tr = sigma_old + C:(eps - eps_old)``,
  SURVEY.md A.2/A.3): active set ``f_trial = seq_tr - sigma_Y(p_old) > 0`` (strict, mfront ``:55``).
* Voce law ``sigma_Y(p) = sig0 + (sigu - sig0)(1 - exp(-b p))`` (``tests/test_FeFp_jax.py:14-15``),
  generalised to ``sig0 + H p + (sigu - sig0)(1 - exp(-b p))`` so linear (``sigu == sig0``), Voce
  (``H == 0``) and heterogeneous batches share one law.  Local solve: scalar Newton on the natural
  residual ``r(dp) = seq_tr - 3 mu dp - sigma_Y(p_old + dp)`` from ``dp = 0`` (monotone: r is convex and
  decreasing), relative tolerance ``|r| <= rtol * seq_tr``, iteration cap -> fail flag.
  The closed form (``n_iter = 0``) is taken when ``b * (sigu - sig0) == 0``.

Every expression below is written component-wise in a fixed operation order (see
``oracle/canon.py``); the CUDA kernel ``dxm_small_strain_kernel`` follows the same order.
"""

import numpy as np

from .canon import exp_c, fma, fnma, lame

NEWTON_CAP = 25
NEWTON_RTOL = 1e-12


def _col(a, n):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 0:
        return np.full(n, float(a))
    return a.reshape(n)


def integrate(eps, state, props, newton_cap=NEWTON_CAP, rtol=NEWTON_RTOL):
    """One batched constitutive update.

    eps   : (n, 6) Mandel strains (C-contiguous float64, as ``quadrature_map.py:313`` builds them)
    state : dict with ``strain`` (n,6), ``stress`` (n,6), ``p`` (n,) or (n,1), ``epsp`` (n,6)  (= s0)
    props : dict with ``E, nu, sig0`` and optionally ``H, sigu, b`` -- scalars or per-point (n,) arrays.
            Pure elasticity: ``sig0 = inf``.  Alternatively ``table = (p_k, sig_k)``: piecewise-linear isotropic
            hardening through the points ``(p_k, sig_k)`` (``p_0 = 0``, increasing), continued with the last slope --
            the device-side stand-in for an arbitrary ``yield_stress`` callable of ``vonMisesIsotropicHardening``;
            the return map walks the segments and is exact (no Newton), ``n_iter`` counts segment crossings.
    Returns a dict: ``strain, stress, p, epsp`` (= s1), ``Ct`` (n,6,6), ``flag`` (n,) uint8 active set,
    ``n_iter`` (n,) int32 local Newton iterations, ``resid`` (n,) final |r|, ``fail`` (n,) uint8.
    """
    eps = np.ascontiguousarray(eps, dtype=np.float64)
    n = eps.shape[0]
    e_old = np.asarray(state["strain"], dtype=np.float64).reshape(n, 6)
    s_old = np.asarray(state["stress"], dtype=np.float64).reshape(n, 6)
    p_old = np.asarray(state["p"], dtype=np.float64).reshape(n)
    ep_old = np.asarray(state["epsp"], dtype=np.float64).reshape(n, 6)

    E = _col(props["E"], n)
    nu = _col(props["nu"], n)
    table = props.get("table")
    if table is not None:
        tp, ts, tH = table_slopes(*table)
        props = dict(props, sig0=ts[0])
    sig0 = _col(props["sig0"], n)
    H = _col(props.get("H", 0.0), n)
    sigu = _col(props.get("sigu", props["sig0"]), n)
    b = _col(props.get("b", 0.0), n)

    with np.errstate(all="ignore"):
        lam, mu = lame(E, nu)
        twomu = 2.0 * mu
        threemu = 3.0 * mu
        dsu = sigu - sig0
        # inf - inf for the elastic class (sig0 = sigu = inf): no hardening term at all
        dsu = np.where(np.isfinite(dsu), dsu, 0.0)
        bdsu = b * dsu

        de = [eps[:, i] - e_old[:, i] for i in range(6)]
        tr = (de[0] + de[1]) + de[2]
        ltr = lam * tr
        st = [s_old[:, i] + fma(twomu, de[i], ltr) for i in range(3)]
        st += [fma(twomu, de[i], s_old[:, i]) for i in range(3, 6)]
        pm = ((st[0] + st[1]) + st[2]) / 3.0
        s = [st[i] - pm for i in range(3)] + [st[i] for i in range(3, 6)]
        ss = s[0] * s[0]
        for i in range(1, 6):
            ss = fma(s[i], s[i], ss)
        seq = np.sqrt(1.5 * ss)

        e0 = exp_c(-(b * p_old))
        sy0 = fma(dsu, 1.0 - e0, fma(H, p_old, sig0))
        if table is not None:
            K = len(tp)
            seg = np.zeros(n, dtype=np.int64)
            for k in range(K - 1):
                seg = np.where(p_old >= tp[k + 1], k + 1, seg)
            sy0 = fma(tH[seg], p_old - tp[seg], ts[seg])
        f = seq - sy0
        flag = f > 0.0

        # ---- local solve ------------------------------------------------------------------
        closed = bdsu == 0.0
        dp = np.zeros(n)
        ecur = e0.copy()
        n_iter = np.zeros(n, dtype=np.int32)
        resid = np.zeros(n)
        fail = np.zeros(n, dtype=bool)

        # closed form (mfront :57-60)
        cf = flag & closed
        dp = np.where(cf, f / (threemu + H), dp)

        if table is not None:
            # exact walk over the segments of the piecewise-linear hardening curve
            closed = np.ones(n, dtype=bool)
            walking = flag.copy()
            for _ in range(K):
                cand = (seq - fma(tH[seg], p_old - tp[seg], ts[seg])) / (threemu + tH[seg])
                dp = np.where(walking, cand, dp)
                nxt = np.minimum(seg + 1, K - 1)
                cross = walking & (seg < K - 1) & (p_old + dp > tp[nxt])
                seg = np.where(cross, seg + 1, seg)
                n_iter = n_iter + cross.astype(np.int32)
                walking = cross
                if not walking.any():
                    break
            H = tH[seg]

        # Newton on the natural residual
        active = flag & ~closed
        tol = rtol * seq
        for it in range(newton_cap + 1):
            if not active.any():
                break
            p = p_old + dp
            sy = fma(dsu, 1.0 - ecur, fma(H, p, sig0))
            r = fnma(threemu, dp, seq) - sy
            conv = np.abs(r) <= tol
            done = active & conv
            resid = np.where(done, np.abs(r), resid)
            active = active & ~conv
            if it == newton_cap:
                fail |= active
                resid = np.where(active, np.abs(r), resid)
                break
            dsy = fma(bdsu, ecur, H)
            dp_new = dp + r / (threemu + dsy)
            dp = np.where(active, dp_new, dp)
            e_new = exp_c(-(b * (p_old + dp)))
            ecur = np.where(active, e_new, ecur)
            n_iter = n_iter + active.astype(np.int32)

        # ---- state update -----------------------------------------------------------------
        Hp = fma(bdsu, ecur, H)  # sigma_Y'(p_new)
        nrm = [np.where(flag, (1.5 * s[i]) / seq, 0.0) for i in range(6)]
        dp = np.where(flag, dp, 0.0)
        depsp = [dp * nrm[i] for i in range(6)]
        sig = [fnma(twomu, depsp[i], st[i]) for i in range(6)]
        epsp = [ep_old[:, i] + depsp[i] for i in range(6)]
        p_new = p_old + dp

        # ---- consistent tangent (mfront :63-66): Ct = A 1x1 + B I6 - gamma n x n ----------------
        q = np.where(flag, dp / seq, 0.0)
        cste = 1.0 / (threemu + Hp)
        fourmu2 = (4.0 * mu) * mu
        beta = fourmu2 * q
        gamma = np.where(flag, fourmu2 * (cste - q), 0.0)
        A = fma(0.5, beta, lam)
        B = fnma(1.5, beta, twomu)
        AB = A + B
        Ct = np.zeros((n, 6, 6))
        for j in range(6):
            for i in range(6):
                if i == j:
                    base = AB if i < 3 else B
                elif i < 3 and j < 3:
                    base = A
                else:
                    base = 0.0
                Ct[:, j, i] = fnma(gamma, nrm[i] * nrm[j], base)

        # fused non-finite check (replaces the host NaN scans of quadrature_map.py:322-324)
        chk = (seq + np.abs(pm)) + p_new
        for i in range(6):
            chk = chk + np.abs(epsp[i])
        fail |= ~np.isfinite(chk)

    return {
        "strain": eps,
        "stress": np.stack(sig, axis=1),
        "p": p_new,
        "epsp": np.stack(epsp, axis=1),
        "Ct": Ct,
        "flag": flag.astype(np.uint8),
        "n_iter": n_iter,
        "resid": resid,
        "fail": fail.astype(np.uint8),
    }


def table_slopes(p, sig):
    """(p_k, sig_k, H_k) of a piecewise-linear hardening table; H_k = slope of segment k, the last one repeated."""
    tp = np.ascontiguousarray(p, dtype=np.float64)
    ts = np.ascontiguousarray(sig, dtype=np.float64)
    if tp.ndim != 1 or tp.size != ts.size or tp.size < 2 or tp[0] != 0.0 or np.any(np.diff(tp) <= 0):
        raise ValueError("hardening table: p must start at 0 and increase strictly, with one stress per point")
    tH = np.empty_like(tp)
    tH[:-1] = (ts[1:] - ts[:-1]) / (tp[1:] - tp[:-1])
    tH[-1] = tH[-2]
    return tp, ts, tH


def zero_state(n):
    """Virgin state, as ``MaterialStateManager.__init__`` zero-initialises it (``generic.py:228-232``)."""
    return {
        "strain": np.zeros((n, 6)),
        "stress": np.zeros((n, 6)),
        "p": np.zeros(n),
        "epsp": np.zeros((n, 6)),
    }


def advance(out):
    """s0 <- s1 (``DataManager.update``, ``generic.py:212-213``)."""
    return {k: out[k] for k in ("strain", "stress", "p", "epsp")}


def elastic_props(E, nu):
    return {"E": E, "nu": nu, "sig0": np.inf}
