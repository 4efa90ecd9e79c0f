"""Frame invariance of the restated behaviours -- properties any correct statement of these isotropic laws has, checked
without a reference (the jaxmat formulations are "parity unpinned"):

* small strain (J2 linear / Voce, Hosford): rotating the whole strain history, eps -> Q eps Q^T, rotates stress and
  plastic strain the same way, leaves p, the active set and the iteration counts' regime unchanged, and transforms the
  tangent as a fourth-order tensor;
* finite strain (FeFp): objectivity under a superposed rigid rotation F -> Q F (PK1 -> Q PK1, be_bar -> Q be_bar Q^T)
  and material isotropy under a change of reference frame F -> F Q^T (PK1 -> PK1 Q^T, be_bar unchanged).
"""
import numpy as np
import pytest

from oracle import fefp
from oracle import hosford as ho
from oracle import small_strain as ss
from oracle import synth

R2 = np.sqrt(2.0)
IDX9 = [(0, 0), (1, 1), (2, 2), (0, 1), (1, 0), (0, 2), (2, 0), (1, 2), (2, 1)]  # utils.py:173-186


def rotation(seed):
    q, r = np.linalg.qr(np.random.default_rng(seed).standard_normal((3, 3)))
    q = q * np.sign(np.diag(r))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q


def mandel_to_tensor(v):
    t = np.empty(v.shape[:-1] + (3, 3))
    t[..., 0, 0], t[..., 1, 1], t[..., 2, 2] = v[..., 0], v[..., 1], v[..., 2]
    t[..., 0, 1] = t[..., 1, 0] = v[..., 3] / R2
    t[..., 0, 2] = t[..., 2, 0] = v[..., 4] / R2
    t[..., 1, 2] = t[..., 2, 1] = v[..., 5] / R2
    return t


def tensor_to_mandel(t):
    return np.stack([t[..., 0, 0], t[..., 1, 1], t[..., 2, 2], R2 * t[..., 0, 1], R2 * t[..., 0, 2], R2 * t[..., 1, 2]], axis=-1)


def mandel_rotation(Q):
    """6x6 orthogonal matrix M with mandel(Q T Q^T) = M mandel(T)."""
    M = np.empty((6, 6))
    for k in range(6):
        e = np.zeros(6)
        e[k] = 1.0
        M[:, k] = tensor_to_mandel(Q @ mandel_to_tensor(e) @ Q.T)
    return M


def v9_to_tensor(v):
    t = np.empty(v.shape[:-1] + (3, 3))
    for k, (i, j) in enumerate(IDX9):
        t[..., i, j] = v[..., k]
    return t


def tensor_to_v9(t):
    return np.stack([t[..., i, j] for (i, j) in IDX9], axis=-1)


SMALL = {
    "j2-linear": (ss.integrate, dict(E=70e3, nu=0.3, sig0=250.0, H=5e3)),
    "j2-voce": (ss.integrate, dict(E=70e3, nu=0.3, sig0=350.0, sigu=500.0, b=1e3)),
    "hosford-10": (ho.integrate, dict(E=70e3, nu=0.3, sig0=200.0, H=10.0, a=10)),
    "hosford-6-voce": (ho.integrate, dict(E=70e3, nu=0.3, sig0=350.0, sigu=500.0, b=1e3, H=25.0, a=6)),
}


@pytest.mark.parametrize("case", list(SMALL))
def test_small_strain_laws_are_isotropic(case):
    integrate, props = SMALL[case]
    n, K = 400, 3
    Q = rotation(1)
    M = mandel_rotation(Q)
    assert np.allclose(M @ M.T, np.eye(6), atol=1e-14)
    st, st_r = ss.zero_state(n), ss.zero_state(n)
    for k in range(1, K + 1):
        eps = synth.strain(n, 2, 1.25e-2, k, K)
        ref = integrate(eps, st, props)
        rot = integrate(np.ascontiguousarray(eps @ M.T), st_r, props)
        assert np.array_equal(ref["flag"], rot["flag"]) and ref["fail"].sum() == 0 and rot["fail"].sum() == 0
        scale = np.abs(ref["stress"]).max()
        assert np.allclose(rot["stress"], ref["stress"] @ M.T, rtol=0, atol=2e-10 * scale)
        assert np.allclose(rot["epsp"], ref["epsp"] @ M.T, rtol=0, atol=1e-10 * max(np.abs(ref["epsp"]).max(), 1e-6))
        assert np.allclose(rot["p"], ref["p"], rtol=1e-9, atol=1e-15)
        # Ct' = M Ct M^T (fourth-order rotation in Mandel form)
        want = np.einsum("ia,nab,jb->nij", M, ref["Ct"], M)
        assert np.allclose(rot["Ct"], want, rtol=0, atol=2e-8 * np.abs(ref["Ct"]).max())
        st, st_r = ss.advance(ref), ss.advance(rot)
    assert 0.3 < ref["flag"].mean() < 1.0


FEFP_PROPS = dict(E=70e3, nu=0.3, sig0=500.0, sigu=750.0, b=1000.0)


def test_fefp_is_objective_under_superposed_rotation():
    """F -> Q_k F with a different rotation per increment: PK1 -> Q_k PK1, be_bar -> Q_k be_bar Q_k^T, p unchanged."""
    n, K = 300, 4
    st, st_r = fefp.virgin_state(n), fefp.virgin_state(n)
    for k in range(1, K + 1):
        Q = rotation(10 + k)
        F = synth.defgrad(n, 4, 3e-2, k, K)
        ref = fefp.integrate(F, st, FEFP_PROPS)
        Fr = tensor_to_v9(Q @ v9_to_tensor(F))
        rot = fefp.integrate(np.ascontiguousarray(Fr), st_r, FEFP_PROPS)
        assert np.array_equal(ref["flag"], rot["flag"]) and rot["fail"].sum() == 0
        scale = np.abs(ref["PK1"]).max()
        assert np.allclose(v9_to_tensor(rot["PK1"]), Q @ v9_to_tensor(ref["PK1"]), rtol=0, atol=1e-9 * scale)
        be, be_r = mandel_to_tensor(ref["be_bar"]), mandel_to_tensor(rot["be_bar"])
        assert np.allclose(be_r, Q @ be @ Q.T, rtol=0, atol=1e-11)
        assert np.allclose(rot["p"], ref["p"], rtol=1e-8, atol=1e-15)
        # dP'_ij / dF'_kl = Q_ia Q_kb dP_aj / dF_bl
        C4 = np.empty((n, 3, 3, 3, 3))
        C4r = np.empty_like(C4)
        for r_, (i, j) in enumerate(IDX9):
            for c_, (kk, l) in enumerate(IDX9):
                C4[:, i, j, kk, l] = ref["Ct"][:, r_, c_]
                C4r[:, i, j, kk, l] = rot["Ct"][:, r_, c_]
        want = np.einsum("ia,kb,najbl->nijkl", Q, Q, C4)
        assert np.allclose(C4r, want, rtol=0, atol=2e-8 * np.abs(C4).max())
        st, st_r = fefp.advance(ref), fefp.advance(rot)
    assert ref["flag"].mean() > 0.3


def test_fefp_is_isotropic_under_change_of_reference_frame():
    """F -> F Q^T (fixed Q for the whole history): PK1 -> PK1 Q^T, spatial state be_bar and p unchanged."""
    n, K = 300, 3
    Q = rotation(3)
    st, st_r = fefp.virgin_state(n), fefp.virgin_state(n)
    st_r["F"] = np.ascontiguousarray(tensor_to_v9(v9_to_tensor(st["F"]) @ Q.T))  # the virgin F = I seen from the new frame
    for k in range(1, K + 1):
        F = synth.defgrad(n, 6, 3e-2, k, K)
        ref = fefp.integrate(F, st, FEFP_PROPS)
        rot = fefp.integrate(np.ascontiguousarray(tensor_to_v9(v9_to_tensor(F) @ Q.T)), st_r, FEFP_PROPS)
        assert np.array_equal(ref["flag"], rot["flag"]) and rot["fail"].sum() == 0
        scale = np.abs(ref["PK1"]).max()
        assert np.allclose(v9_to_tensor(rot["PK1"]), v9_to_tensor(ref["PK1"]) @ Q.T, rtol=0, atol=1e-9 * scale)
        assert np.allclose(rot["be_bar"], ref["be_bar"], rtol=0, atol=1e-11)
        assert np.allclose(rot["p"], ref["p"], rtol=1e-8, atol=1e-15)
        st, st_r = fefp.advance(ref), fefp.advance(rot)
    assert ref["flag"].mean() > 0.3


def test_fefp_small_strain_limit_is_the_j2_voce_law():
    """Two independently formulated restatements meet: for displacement gradients of order 1e-5 (and a yield stress scaled
    down so that most points still yield) the multiplicative FeFp update reduces to the additive small-strain J2 + Voce
    update to first order in |grad u| -- PK1 -> sigma, same cumulated plastic strain, dP/dF -> the small-strain tangent."""
    n, K, amp = 1500, 3, 4e-5
    props = dict(E=70e3, nu=0.3, sig0=0.5, sigu=0.75, b=2e5)
    st9, st6 = fefp.virgin_state(n), ss.zero_state(n)
    for k in range(1, K + 1):
        F = synth.defgrad(n, 8, amp, k, K)
        H = v9_to_tensor(F) - np.eye(3)
        eps = tensor_to_mandel(0.5 * (H + np.swapaxes(H, -1, -2)))
        fin = fefp.integrate(F, st9, props)
        sml = ss.integrate(np.ascontiguousarray(eps), st6, props)
        small = np.abs(H).max()  # ~1e-4: size of the neglected geometric terms
        assert small < 2e-4 and fin["fail"].sum() == 0 and sml["fail"].sum() == 0
        sig = mandel_to_tensor(sml["stress"])
        P = v9_to_tensor(fin["PK1"])
        assert np.abs(P - sig).max() <= 5 * small * np.abs(sig).max()  # measured: 1.7 |grad u|
        both = (fin["flag"] == 1) & (sml["flag"] == 1)
        assert (fin["flag"] != sml["flag"]).mean() < 0.01  # only points within O(|grad u|) of the yield surface may differ
        assert np.allclose(fin["p"][both], sml["p"][both], rtol=50 * small, atol=1e-4 * sml["p"].max())
        # dP_ij/dF_kl vs C_ijkl of the small-strain tangent (minor symmetries: Mandel factors undone)
        C4 = np.empty((n, 3, 3, 3, 3))
        for r_, (i, j) in enumerate(IDX9):
            for c_, (kk, l) in enumerate(IDX9):
                C4[:, i, j, kk, l] = fin["Ct"][:, r_, c_]
        M6 = [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]
        C4s = np.empty_like(C4)
        for a_, (i, j) in enumerate(M6):
            for b_, (kk, l) in enumerate(M6):
                v = sml["Ct"][:, a_, b_] / ((R2 if i != j else 1.0) * (R2 if kk != l else 1.0))
                for (ii, jj) in {(i, j), (j, i)}:
                    for (k2, l2) in {(kk, l), (l, kk)}:
                        C4s[:, ii, jj, k2, l2] = v
        agree = fin["flag"] == sml["flag"]
        err = np.abs(C4[agree] - C4s[agree]).reshape(agree.sum(), -1).max(axis=1)
        # first order in |grad u| almost everywhere; points sitting within O(|grad u|) of the yield surface have an
        # O(1) sensitivity of the plastic moduli to that perturbation
        assert np.quantile(err, 0.98) <= 50 * small * np.abs(C4s).max()
        st9, st6 = fefp.advance(fin), ss.advance(sml)
    assert both.mean() > 0.3
