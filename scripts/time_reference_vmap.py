"""SURVEY 8(d), CPU baseline (1): the reference's OWN code for this path where it can run -- `generic.Material.integrate`
with its Python-loop vectoriser `_vmap` (generic.py:10-100, :176-189) driving (a) the in-tree
`LinearElasticIsotropic` (python_materials/elasticity.py:5-24) and (b) a per-point J2+Voce `constitutive_update` that
calls the oracle for one point (the shape a pure-Python plasticity material takes in the reference protocol).
Build container only (imports /root/reference; the GPU boxes do not have it):
    python scripts/time_reference_vmap.py  ->  profiles/r01_reference_vmap_cpu.json"""
import json, os, sys, time, types, warnings
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dolfinx = types.ModuleType("dolfinx"); common = types.ModuleType("dolfinx.common")
class Timer:
    def __init__(self, *a, **k): pass
    def __enter__(self): return self
    def __exit__(self, *a): return False
common.Timer = Timer; dolfinx.common = common
sys.modules["dolfinx"] = dolfinx; sys.modules["dolfinx.common"] = common
sys.path.insert(0, "/root/reference")
from dolfinx_materials.generic import Material  # the reference
from dolfinx_materials.python_materials.elasticity import LinearElasticIsotropic
from oracle import small_strain as ss, synth
warnings.simplefilter("ignore")

PROPS = dict(E=70e3, nu=0.3, sig0=350.0, sigu=500.0, b=1e3)


class PointwiseJ2(Material):
    """A reference-protocol material whose per-point update is one-point J2+Voce (state p, epsp)."""
    @property
    def internal_state_variables(self):
        return {"p": 1, "epsp": 6}

    def constitutive_update(self, eps, state, dt):
        st = {"strain": state["Strain"][None, :], "stress": state["Stress"][None, :], "p": state["p"], "epsp": state["epsp"][None, :]}
        out = ss.integrate(eps[None, :], st, PROPS)
        state["Strain"], state["Stress"], state["p"], state["epsp"] = eps, out["stress"][0], out["p"], out["epsp"][0]
        return out["Ct"][0], state


res = []
for name, mat, n in (("LinearElasticIsotropic (in-tree)", LinearElasticIsotropic(E=70e3, nu=0.3), 100_000),
                     ("per-point J2+Voce through the reference protocol", PointwiseJ2(), 5_000)):
    mat.set_data_manager(n)
    eps = synth.strain(n, 0, 1.25e-2, 1, 1)
    t0 = time.perf_counter(); mat.integrate(eps); dt = time.perf_counter() - t0
    res.append(dict(material=name, n=n, seconds=dt, points_per_s=n / dt, cores=1,
                    path="dolfinx_materials.generic.Material.integrate -> _vmap (Python loop over Gauss points)"))
    print(res[-1])
json.dump(res, open(os.path.join(ROOT, "profiles", "r01_reference_vmap_cpu.json"), "w"), indent=1)
