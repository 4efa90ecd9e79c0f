"""One build of the library (default: hand-placed fused multiply-adds; DXM_UNFUSED=1: every fma split again = the
round-1 arithmetic) against the committed golden histories and the clock: prints one JSON line with, per history,
whether the results are bit-identical to the fixture and their worst relative deviation from it, and the kernel time
of the FeFp and Hosford updates on seeded batches.  tests/test_unfused_gpu.py runs it under both builds."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dolfinx_materials_b200 as jm
from dolfinx_materials_b200 import build

build.build_library()
out = dict(unfused=os.environ.get("DXM_UNFUSED", "0"), histories={})
el = jm.LinearElasticIsotropic(E=70e3, nu=0.3)
for name in ("j2_voce_history", "j2_linear_history", "fefp_history", "hosford_history"):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    p = dict(zip([str(k) for k in g["props_keys"]], [float(v) for v in g["props_vals"]]))
    el_h = jm.LinearElasticIsotropic(E=p["E"], nu=p["nu"])
    hard = jm.VoceHardening(sig0=p["sig0"], sigu=p["sigu"], b=p["b"]) if "b" in p else jm.LinearHardening(sig0=p["sig0"], H=p["H"])
    finite = name.startswith("fefp")
    if finite:
        beh = jm.FeFpJ2Plasticity(elasticity=el_h, yield_stress=hard)
    elif name.startswith("hosford"):
        beh = jm.GeneralIsotropicHardening(elasticity=el_h, yield_stress=hard, equivalent_stress=jm.Hosford(a=int(p["a"])))
    else:
        beh = jm.vonMisesIsotropicHardening(elasticity=el_h, yield_stress=hard)
    key = "F" if finite else "eps"
    m = jm.CUDAMaterial(beh)
    m.set_data_manager(g[f"{key}1"].shape[0])
    k, same, dev = 1, True, 0.0
    while f"{key}{k}" in g:
        flux, isv, Ct = m.integrate(g[f"{key}{k}"])
        for got, want in ((flux, g[f"flux{k}"]), (isv, g[f"isv{k}"]), (Ct, g[f"Ct{k}"])):
            same = same and bool(np.array_equal(got, want))
            dev = max(dev, float(np.abs(np.asarray(got) - want).max() / np.abs(want).max()))
        m.data_manager.update()
        k += 1
    out["histories"][name] = dict(bit_identical=same, max_rel_dev=dev)
# throughput A/B on seeded batches
n = 4_000_000
mf = jm.CUDAMaterial(jm.FeFpJ2Plasticity(elasticity=el, yield_stress=jm.VoceHardening(sig0=500.0, sigu=750.0, b=1000.0)))
mf.set_data_manager(n)
for k in range(1, 4):
    mf.synth_gradients(0, 3e-2, k, 4); mf.integrate_resident(); mf.data_manager.update()
mf.synth_gradients(0, 3e-2, 4, 4)
out["fefp_ms"] = sorted(mf.integrate_resident().kernel_ms for _ in range(9))[4]
del mf
mh = jm.CUDAMaterial(jm.GeneralIsotropicHardening(elasticity=el, yield_stress=jm.LinearHardening(sig0=200.0, H=10.0)))
mh.set_data_manager(n)
mh.synth_gradients(0, 1.25e-2, 1, 1)
out["hosford_ms"] = sorted(mh.integrate_resident().kernel_ms for _ in range(9))[4]
out["n"] = n
print(json.dumps(out))
