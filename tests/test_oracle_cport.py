"""The plain-C oracle and the numpy oracle are two independent CPU restatements of the same canonical
arithmetic: they must agree bit for bit (flags, iteration counts, stress, state, tangent)."""
import numpy as np

from oracle import cport, fefp, synth
from oracle import small_strain as ss

KEYS_SS = ("stress", "p", "epsp", "Ct", "flag", "n_iter", "resid", "fail")
KEYS_FE = ("PK1", "p", "be_bar", "Ct", "flag", "n_iter", "resid", "fail")


def test_small_strain_c_equals_numpy():
    n, K = 20011, 4
    for props in (dict(E=70e3, nu=0.3, sig0=350.0, sigu=500.0, b=1e3), dict(E=70e3, nu=0.3, sig0=250.0, H=5e3),
                  ss.elastic_props(70e3, 0.3)):
        st = ss.zero_state(n)
        for k in range(1, K + 1):
            eps = synth.strain(n, 0, 1.25e-2, k, K)
            a, b = ss.integrate(eps, st, props), cport.small_strain(eps, st, props)
            for key in KEYS_SS:
                assert np.array_equal(a[key], b[key]), key
            st = ss.advance(a)


def test_small_strain_c_per_point_properties_and_failures():
    n = 3000
    cls = np.arange(n) % 3
    props = {"E": np.where(cls == 1, 90e3, 70e3), "nu": np.where(cls == 1, 0.25, 0.3), "sig0": np.where(cls == 2, np.inf, 200.0),
             "H": np.where(cls == 0, 10.0, 0.0), "sigu": np.where(cls == 1, 300.0, np.where(cls == 2, np.inf, 200.0)),
             "b": np.where(cls == 1, 10.0, 0.0)}
    eps = synth.strain(n, 9, 1.25e-2, 1, 1)
    eps[5, 1] = np.nan
    a, b = ss.integrate(eps, ss.zero_state(n), props), cport.small_strain(eps, ss.zero_state(n), props)
    for key in KEYS_SS:
        assert np.array_equal(a[key], b[key], equal_nan=(key in ("stress", "epsp", "Ct", "p", "resid"))), key
    a = ss.integrate(eps[10:11], ss.zero_state(1), dict(E=70e3, nu=0.3, sig0=100.0, sigu=500.0, b=1e3), newton_cap=1)
    b = cport.small_strain(eps[10:11], ss.zero_state(1), dict(E=70e3, nu=0.3, sig0=100.0, sigu=500.0, b=1e3), newton_cap=1)
    assert a["fail"][0] == b["fail"][0] and a["n_iter"][0] == b["n_iter"][0]


def test_fefp_c_equals_numpy():
    n, K = 5003, 4
    props = dict(E=70e3, nu=0.3, sig0=500.0, sigu=750.0, b=1000.0)
    st = fefp.virgin_state(n)
    for k in range(1, K + 1):
        F = synth.defgrad(n, 0, 4e-2, k, K)
        a, b = fefp.integrate(F, st, props), cport.fefp(F, st, props)
        for key in KEYS_FE:
            assert np.array_equal(a[key], b[key]), key
        st = fefp.advance(a)
    assert a["flag"].mean() > 0.3
