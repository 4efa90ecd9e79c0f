"""Generates the committed golden fixtures by running the REFERENCE's own code where it can run.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py

What the reference itself computes here (imported from /root/reference with a stub for its single,
unused dolfinx import `dolfinx.common.Timer`, generic.py:2):
  * `dolfinx_materials.python_materials.elasticity.LinearElasticIsotropic` driven through
    `generic.Material.integrate` / `_vmap` / `DataManager` / `MaterialStateManager`
    -> elastic_reference.npz  (flux, Ct, state after update) : pins the elastic update AND the s0/s1
       state machinery.
  * the same reference machinery (`Material.integrate`, per-point `_vmap` loop, `s1.set_item`,
    `data_manager.update()`) driving a per-point `constitutive_update` that calls the oracle for one
    point -> j2_voce_history.npz / j2_linear_history.npz : pins that the batched oracle + state carry
    equals the reference's point-by-point protocol over a load history; fefp_history.npz likewise for the
    finite-strain behaviour (gradient F, flux PK1, state p + be_bar).  (The J2 arithmetic itself is
    not in the reference tree -- it lives in un-vendored jaxmat -- so these vectors pin protocol and
    regression, not jaxmat parity; see oracle/__init__.py.)  hosford_history.npz: the same for the Hosford behaviour
    (oracle/hosford.py; MFront parity unpinned).
"""
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

# --- stub for `from dolfinx.common import Timer` (imported, never used, by generic.py) ------------
dolfinx = types.ModuleType("dolfinx")
common = types.ModuleType("dolfinx.common")


class Timer:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


common.Timer = Timer
dolfinx.common = common
sys.modules["dolfinx"] = dolfinx
sys.modules["dolfinx.common"] = common
sys.path.insert(0, "/root/reference")

from dolfinx_materials.generic import Material  # noqa: E402  (the reference)
from dolfinx_materials.python_materials.elasticity import LinearElasticIsotropic  # noqa: E402

from oracle import fefp  # noqa: E402
from oracle import hosford as ho  # noqa: E402
from oracle import small_strain as ss  # noqa: E402
from oracle import synth  # noqa: E402

warnings.simplefilter("ignore")


def elastic_reference():
    n = 64
    mat = LinearElasticIsotropic(E=70e3, nu=0.3)
    mat.set_data_manager(n)
    out = {}
    for k in (1, 2):
        eps = synth.strain(n, 11, 2e-2, k, 2)
        flux, isv, Ct = mat.integrate(eps)
        out[f"eps{k}"] = eps
        out[f"flux{k}"] = np.array(flux)
        out[f"Ct{k}"] = np.array(Ct)
        mat.data_manager.update()
        out[f"s0_stress_after_update{k}"] = np.array(mat.get_initial_state_dict()["Stress"])
    # nu = 0: sigma = E eps  (tests/mfront/test_initialization.py:131-153)
    mat0 = LinearElasticIsotropic(E=70e3, nu=0.0)
    mat0.set_data_manager(1)
    flux, _, _ = mat0.integrate(np.array([[1e-3, 0, 0, 0, 0, 0.0]]))
    out["flux_nu0"] = np.array(flux)
    np.savez(os.path.join(HERE, "elastic_reference.npz"), **out)


class PointwiseJ2(Material):
    """A per-point material in the reference's own style (constitutive_update(eps, state, dt) ->
    (Ct, state)), so that the REFERENCE's integrate/_vmap/DataManager code drives the history."""

    def __init__(self, props):
        super().__init__()
        self.props = props

    @property
    def gradients(self):
        return {"strain": 6}

    @property
    def fluxes(self):
        return {"stress": 6}

    @property
    def internal_state_variables(self):
        return {"p": 1, "epsp": 6}

    def constitutive_update(self, eps, state, dt):
        st = {k: np.asarray(v, dtype=float).reshape(1, -1) for k, v in state.items()}
        st["p"] = st["p"].reshape(1)
        out = ss.integrate(eps.reshape(1, 6), st, self.props)
        new = {"strain": eps, "stress": out["stress"][0], "p": out["p"], "epsp": out["epsp"][0]}
        return out["Ct"][0], new


def j2_history(name, props, n, amp, K, seed):
    mat = PointwiseJ2(props)
    mat.set_data_manager(n)
    out = {"props_keys": np.array(sorted(props)), "props_vals": np.array([props[k] for k in sorted(props)])}
    for k in range(1, K + 1):
        eps = synth.strain(n, seed, amp, k, K)
        flux, isv, Ct = mat.integrate(eps)
        out[f"eps{k}"] = eps
        out[f"flux{k}"] = np.array(flux)
        out[f"isv{k}"] = np.array(isv)
        out[f"Ct{k}"] = np.array(Ct)
        mat.data_manager.update()
    np.savez(os.path.join(HERE, name), **out)


class PointwiseFeFp(Material):
    """FeFp as a per-point reference-style material: the reference's own integrate/_vmap/DataManager drive it."""

    def __init__(self, props):
        super().__init__()
        self.props = props

    @property
    def gradients(self):
        return {"F": 9}

    @property
    def fluxes(self):
        return {"PK1": 9}

    @property
    def internal_state_variables(self):
        return {"p": 1, "be_bar": 6}

    def constitutive_update(self, F, state, dt):
        st = {"F": state["F"].reshape(1, 9), "PK1": state["PK1"].reshape(1, 9), "p": state["p"].reshape(1),
              "be_bar": state["be_bar"].reshape(1, 6)}
        out = fefp.integrate(F.reshape(1, 9), st, self.props)
        new = {"F": F, "PK1": out["PK1"][0], "p": out["p"], "be_bar": out["be_bar"][0]}
        return out["Ct"][0], new


def fefp_history(name, props, n, amp, K, seed):
    mat = PointwiseFeFp(props)
    mat.set_data_manager(n)
    virgin = fefp.virgin_state(n)
    mat.set_initial_state_dict({"F": virgin["F"], "be_bar": virgin["be_bar"]})
    out = {"props_keys": np.array(sorted(props)), "props_vals": np.array([props[k] for k in sorted(props)])}
    for k in range(1, K + 1):
        F = synth.defgrad(n, seed, amp, k, K)
        flux, isv, Ct = mat.integrate(F)
        out[f"F{k}"] = F
        out[f"flux{k}"] = np.array(flux)
        out[f"isv{k}"] = np.array(isv)
        out[f"Ct{k}"] = np.array(Ct)
        mat.data_manager.update()
    np.savez(os.path.join(HERE, name), **out)


class PointwiseHosford(PointwiseJ2):
    """Hosford criterion + linear hardening, point by point through the reference's machinery (the behaviour the
    multi-material demo takes from MFront, demos/multimaterials/multimaterials.py:245-254)."""

    def constitutive_update(self, eps, state, dt):
        st = {k: np.asarray(v, dtype=float).reshape(1, -1) for k, v in state.items()}
        st["p"] = st["p"].reshape(1)
        out = ho.integrate(eps.reshape(1, 6), st, self.props)
        new = {"strain": eps, "stress": out["stress"][0], "p": out["p"], "epsp": out["epsp"][0]}
        return out["Ct"][0], new


def hosford_history(name, props, n, amp, K, seed):
    mat = PointwiseHosford(props)
    mat.set_data_manager(n)
    out = {"props_keys": np.array(sorted(props)), "props_vals": np.array([float(props[k]) for k in sorted(props)])}
    for k in range(1, K + 1):
        eps = synth.strain(n, seed, amp, k, K)
        flux, isv, Ct = mat.integrate(eps)
        out[f"eps{k}"] = eps
        out[f"flux{k}"] = np.array(flux)
        out[f"isv{k}"] = np.array(isv)
        out[f"Ct{k}"] = np.array(Ct)
        mat.data_manager.update()
    np.savez(os.path.join(HERE, name), **out)


if __name__ == "__main__":
    elastic_reference()
    j2_history("j2_voce_history.npz", dict(E=70e3, nu=0.3, sig0=350.0, sigu=500.0, b=1e3), 96, 1.25e-2, 4, 0)
    j2_history("j2_linear_history.npz", dict(E=70e3, nu=0.3, sig0=250.0, H=5e3), 48, 1.25e-2, 3, 5)
    hosford_history("hosford_history.npz", dict(E=70e3, nu=0.3, sig0=200.0, H=10.0, a=10), 48, 1.25e-2, 3, 7)
    fefp_history("fefp_history.npz", dict(E=70e3, nu=0.3, sig0=500.0, sigu=750.0, b=1000.0), 40, 4e-2, 3, 3)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
