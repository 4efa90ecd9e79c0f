"""Fixtures of EXACT solutions: the reference behaviours' own equations solved in 40-digit arithmetic
(``oracle/jaxmat_form_mp.py``, ``oracle/hosford_mp.py``), rounded to double, for seeded two-increment histories.

    python tests/golden/make_golden_exact.py        # writes tests/golden/exact_{j2_voce,fefp,hosford}.npz

``tests/test_golden_exact.py`` holds the CPU oracle and -- on a B200 -- ``CUDAMaterial`` to them: stress within 2e-12,
plastic multiplier within the local Newton tolerance, tangent within 1e-11.  The state handed from increment 1 to 2 is the
exact one (rounded), so the fixtures do not depend on any double-precision implementation."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import fefp, hosford as ho, hosford_mp as hm, jaxmat_form_mp as jmp, small_strain as ss, synth  # noqa: E402

VOCE = dict(E=70e3, nu=0.3, sig0=350.0, sigu=500.0, b=1e3)
FEFP = dict(E=70e3, nu=0.3, sig0=500.0, sigu=750.0, b=1000.0)
HOSF = dict(E=70e3, nu=0.3, sig0=200.0, H=10.0, a=10)
f = lambda v: np.array([float(x) for x in v])  # noqa: E731
mat = lambda M, n: np.array([[float(M[r, c]) for c in range(n)] for r in range(n)])  # noqa: E731


def j2(n=12):
    st = ss.zero_state(n)
    rec = {k: [] for k in ("eps", "stress", "p", "epsp", "Ct")}
    for k in (1, 2):
        eps = synth.strain(n, 11, 1.25e-2, k, 2)
        guess = ss.integrate(eps, st, VOCE)  # starting values for the 40-digit root finder only
        S, P, EP, CT = [], [], [], []
        for i in range(n):
            dp0 = float(guess["p"][i] - st["p"][i])
            r = jmp.j2_point(eps[i], st["strain"][i], st["stress"][i], st["p"][i], VOCE, dp0=dp0)
            S.append(f(r["stress"])); P.append(float(r["p"])); EP.append(st["epsp"][i] + f(r["depsp"]))
            CT.append(mat(jmp.j2_tangent(eps[i], st["strain"][i], st["stress"][i], st["p"][i], VOCE, dp0=dp0), 6))
        st = {"strain": eps.copy(), "stress": np.array(S), "p": np.array(P), "epsp": np.array(EP)}
        for key, val in (("eps", eps), ("stress", S), ("p", P), ("epsp", EP), ("Ct", CT)):
            rec[key].append(np.array(val))
    np.savez_compressed(os.path.join(HERE, "exact_j2_voce.npz"), props=np.array(list(VOCE.items()), dtype=object),
                        **{k: np.stack(v) for k, v in rec.items()})


def finite(n=8):
    st = fefp.virgin_state(n)
    rec = {k: [] for k in ("F", "PK1", "p", "be_bar", "Ct")}
    for k in (1, 2):
        F = synth.defgrad(n, 11, 3e-2, k, 2)
        guess = fefp.integrate(F, st, FEFP)
        Pk, P, BE, CT = [], [], [], []
        for i in range(n):
            start = (float(guess["p"][i] - st["p"][i]), guess["be_bar"][i])
            r = jmp.fefp_point(F[i], st["F"][i], st["be_bar"][i], st["p"][i], FEFP, start=start)
            Pk.append(f(r["PK1"])); P.append(float(r["p"])); BE.append(f(r["be_bar"]))
            CT.append(mat(jmp.fefp_tangent(F[i], st["F"][i], st["be_bar"][i], st["p"][i], FEFP, start=start), 9))
        st = dict(st, F=F.copy(), PK1=np.array(Pk), p=np.array(P), be_bar=np.array(BE))
        for key, val in (("F", F), ("PK1", Pk), ("p", P), ("be_bar", BE), ("Ct", CT)):
            rec[key].append(np.array(val))
    np.savez_compressed(os.path.join(HERE, "exact_fefp.npz"), props=np.array(list(FEFP.items()), dtype=object),
                        **{k: np.stack(v) for k, v in rec.items()})


def hosford(n=8):
    st = ss.zero_state(n)
    rec = {k: [] for k in ("eps", "stress", "p", "epsp", "Ct")}
    for k in (1, 2):
        eps = synth.strain(n, 11, 1.0e-2, k, 2)
        if k == 1:
            eps[0, 1:] = 0.0  # uniaxial strain: a repeated eigenvalue
        guess = ho.integrate(eps, st, HOSF)
        S, P, EP, CT = [], [], [], []
        for i in range(n):
            d_eel = (eps[i] - st["strain"][i]) - (guess["epsp"][i] - st["epsp"][i])
            start = (list(d_eel), float(guess["p"][i] - st["p"][i])) if guess["flag"][i] else None
            r = hm.integrate_point(eps[i], st["strain"][i], st["epsp"][i], st["p"][i], HOSF, start=start)
            S.append(f(r["stress"])); P.append(float(r["p"])); EP.append(f(r["epsp"]))
            CT.append(mat(hm.tangent_point(eps[i], st["strain"][i], st["epsp"][i], st["p"][i], HOSF, start=start), 6))
        st = {"strain": eps.copy(), "stress": np.array(S), "p": np.array(P), "epsp": np.array(EP)}
        for key, val in (("eps", eps), ("stress", S), ("p", P), ("epsp", EP), ("Ct", CT)):
            rec[key].append(np.array(val))
    np.savez_compressed(os.path.join(HERE, "exact_hosford.npz"), props=np.array(list(HOSF.items()), dtype=object),
                        **{k: np.stack(v) for k, v in rec.items()})


if __name__ == "__main__":
    j2(); finite(); hosford()
    for name in ("exact_j2_voce", "exact_fefp", "exact_hosford"):
        d = np.load(os.path.join(HERE, name + ".npz"), allow_pickle=True)
        print(name, {k: d[k].shape for k in d.files if k != "props"})
