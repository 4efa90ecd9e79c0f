"""The CPU oracle and -- on a B200 -- ``CUDAMaterial`` against committed fixtures of EXACT solutions: the reference
behaviours' own equations (jaxmat's Fischer-Burmeister systems, the seven-unknown system of MFront's ``Implicit`` DSL for
Hosford) solved in 40-digit arithmetic and rounded to double (``tests/golden/make_golden_exact.py``,
``tests/golden/exact_*.npz``).  Two increments; the state handed to the second one is the exact one, so nothing in the
fixtures depends on a double-precision implementation.  Bars (written out, all inside the north star's rtol 1e-10):
stress 2e-12 of the largest component, plastic multiplier / internal state 1e-11 (the local Newton stops at 1e-12 of the
equivalent stress), consistent tangent 1e-11 of its largest entry."""

import os

import numpy as np
import pytest

from oracle import fefp
from oracle import hosford as ho
from oracle import small_strain as ss

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL_STRESS, RTOL_STATE, RTOL_TANGENT = 2e-12, 1e-11, 1e-11


def _load(name):
    d = np.load(os.path.join(GOLD, name), allow_pickle=True)
    props = {k: (int(v) if k == "a" else float(v)) for k, v in d["props"]}
    return d, props


def _close(got, ref, rtol):
    scale = np.abs(ref).max(axis=tuple(range(1, ref.ndim)), keepdims=True) if ref.ndim > 1 else np.abs(ref)
    return np.all(np.abs(got - ref) <= rtol * np.maximum(scale, 1e-300) + 1e-18)


def _small_strain_states(d, k):
    n = d["eps"].shape[1]
    if k == 0:
        return ss.zero_state(n)
    return {"strain": d["eps"][k - 1], "stress": d["stress"][k - 1], "p": d["p"][k - 1], "epsp": d["epsp"][k - 1]}


def _check_small(out_stress, out_p, out_epsp, out_ct, d, k):
    assert _close(out_stress, d["stress"][k], RTOL_STRESS)
    assert _close(out_p, d["p"][k], RTOL_STATE)
    assert np.abs(out_epsp - d["epsp"][k]).max() <= RTOL_STATE * np.abs(d["eps"][k]).max()
    assert _close(out_ct.reshape(-1, 36), d["Ct"][k].reshape(-1, 36), RTOL_TANGENT)


@pytest.mark.parametrize("name, integrate", [("exact_j2_voce.npz", ss.integrate), ("exact_hosford.npz", ho.integrate)])
def test_oracle_small_strain_vs_exact_fixtures(name, integrate):
    d, props = _load(name)
    plastic = 0
    for k in range(2):
        out = integrate(d["eps"][k], _small_strain_states(d, k), props)
        _check_small(out["stress"], out["p"], out["epsp"], out["Ct"], d, k)
        plastic += int(out["flag"].sum())
    assert plastic >= d["eps"].shape[1]


def test_oracle_finite_strain_vs_exact_fixture():
    d, props = _load("exact_fefp.npz")
    n = d["F"].shape[1]
    st = fefp.virgin_state(n)
    for k in range(2):
        out = fefp.integrate(d["F"][k], st, props)
        assert _close(out["PK1"], d["PK1"][k], 5e-12) and _close(out["p"], d["p"][k], RTOL_STATE)
        assert np.abs(out["be_bar"] - d["be_bar"][k]).max() <= 1e-12
        assert _close(out["Ct"].reshape(n, 81), d["Ct"][k].reshape(n, 81), RTOL_TANGENT)
        st = dict(st, F=d["F"][k], PK1=d["PK1"][k], p=d["p"][k], be_bar=d["be_bar"][k])


# ---- the same on the GPU, through the public API ---------------------------------------------------------------------
def _material(jm, name, props):
    el = jm.LinearElasticIsotropic(E=props["E"], nu=props["nu"])
    if name == "exact_j2_voce.npz":
        return jm.CUDAMaterial(jm.vonMisesIsotropicHardening(
            elasticity=el, yield_stress=jm.VoceHardening(sig0=props["sig0"], sigu=props["sigu"], b=props["b"])))
    if name == "exact_hosford.npz":
        return jm.CUDAMaterial(jm.GeneralIsotropicHardening(
            elasticity=el, yield_stress=jm.LinearHardening(sig0=props["sig0"], H=props["H"]), equivalent_stress=jm.Hosford(a=props["a"])))
    return jm.CUDAMaterial(jm.FeFpJ2Plasticity(
        elasticity=el, yield_stress=jm.VoceHardening(sig0=props["sig0"], sigu=props["sigu"], b=props["b"])))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["exact_j2_voce.npz", "exact_hosford.npz"])
def test_gpu_small_strain_vs_exact_fixtures(jm, name):
    d, props = _load(name)
    n = d["eps"].shape[1]
    m = _material(jm, name, props)
    m.set_data_manager(n)
    for k in range(2):
        if k:
            m.set_initial_state_dict({"strain": d["eps"][0], "stress": d["stress"][0], "p": d["p"][0], "epsp": d["epsp"][0]})
        flux, isv, Ct = m.integrate(d["eps"][k])
        _check_small(flux, isv[:, 0], isv[:, 1:], Ct, d, k)
        assert m.last_stats.n_fail == 0


@pytest.mark.gpu
def test_gpu_finite_strain_vs_exact_fixture(jm):
    d, props = _load("exact_fefp.npz")
    n = d["F"].shape[1]
    m = _material(jm, "exact_fefp.npz", props)
    m.set_data_manager(n)
    for k in range(2):
        if k:
            m.set_initial_state_dict({"F": d["F"][0], "PK1": d["PK1"][0], "p": d["p"][0], "be_bar": d["be_bar"][0]})
        flux, isv, Ct = m.integrate(d["F"][k])
        assert _close(flux, d["PK1"][k], 5e-12) and _close(isv[:, 0], d["p"][k], RTOL_STATE)
        assert np.abs(isv[:, 1:] - d["be_bar"][k]).max() <= 1e-12
        assert _close(Ct.reshape(n, 81), d["Ct"][k].reshape(n, 81), RTOL_TANGENT)
