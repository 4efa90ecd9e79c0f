"""DXM_FMAD=1 build: deviation from the canonical oracle (it is no longer bit-identical) and throughput."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dolfinx_materials_b200 as jm
from dolfinx_materials_b200 import build
from oracle import fefp, hosford as ho, small_strain as ss, synth
build.build_library()
el = jm.LinearElasticIsotropic(E=70e3, nu=0.3)
out = dict(fmad=os.environ.get("DXM_FMAD", "0"))
n = 200_000
# J2 + Voce
V = dict(E=70e3, nu=0.3, sig0=350.0, sigu=500.0, b=1e3)
m = jm.CUDAMaterial(jm.vonMisesIsotropicHardening(elasticity=el, yield_stress=jm.VoceHardening(sig0=350.0, sigu=500.0, b=1e3)))
m.set_data_manager(n); m.enable_diagnostics(); st = ss.zero_state(n)
for k in range(1, 5):
    eps = synth.strain(n, 0, 1.25e-2, k, 4); flux, isv, Ct = m.integrate(eps); ref = ss.integrate(eps, st, V)
    flag, n_iter, _, _ = m.diagnostics(); m.data_manager.update(); st = ss.advance(ref)
rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
out["j2"] = dict(stress=rel(flux, ref["stress"]), Ct=rel(Ct, ref["Ct"]), p=rel(isv[:, 0], ref["p"]),
                 flag_mismatch=int((flag != ref["flag"]).sum()), iter_mismatch=int((n_iter != ref["n_iter"]).sum()))
F = dict(E=70e3, nu=0.3, sig0=500.0, sigu=750.0, b=1000.0)
fm = jm.CUDAMaterial(jm.FeFpJ2Plasticity(elasticity=el, yield_stress=jm.VoceHardening(sig0=500.0, sigu=750.0, b=1000.0)))
fm.set_data_manager(n); fm.enable_diagnostics(); fst = fefp.virgin_state(n)
for k in range(1, 5):
    G = synth.defgrad(n, 0, 3e-2, k, 4); P, isv, Ct = fm.integrate(G); ref = fefp.integrate(G, fst, F)
    flag, n_iter, _, _ = fm.diagnostics(); fm.data_manager.update(); fst = fefp.advance(ref)
out["fefp"] = dict(PK1=rel(P, ref["PK1"]), Ct=rel(Ct, ref["Ct"]), p=rel(isv[:, 0], ref["p"]),
                   flag_mismatch=int((flag != ref["flag"]).sum()), iter_mismatch=int((n_iter != ref["n_iter"]).sum()))
# Hosford (a = 10): the local Newton's line search compares merit values, so a handful of iteration counts may move
H = dict(E=70e3, nu=0.3, sig0=200.0, H=10.0, a=10)
hm = jm.CUDAMaterial(jm.GeneralIsotropicHardening(elasticity=el, yield_stress=jm.LinearHardening(sig0=200.0, H=10.0)))
hm.set_data_manager(n); hm.enable_diagnostics(); hst = ss.zero_state(n)
for k in range(1, 5):
    eps = synth.strain(n, 0, 1.25e-2, k, 4); flux, isv, Ct = hm.integrate(eps); ref = ho.integrate(eps, hst, H)
    flag, n_iter, _, _ = hm.diagnostics(); hm.data_manager.update(); hst = ss.advance(ref)
out["hosford"] = dict(stress=rel(flux, ref["stress"]), Ct=rel(Ct, ref["Ct"]), p=rel(isv[:, 0], ref["p"]),
                      flag_mismatch=int((flag != ref["flag"]).sum()), iter_mismatch=int((n_iter != ref["n_iter"]).sum()))
n2 = 4_000_000
hb = jm.CUDAMaterial(hm.behavior); hb.set_data_manager(n2); hb.synth_gradients(0, 1.25e-2, 1, 1)
os.environ["DXM_HOS_SPLIT"] = "0"
ts = sorted(hb.integrate_resident().kernel_ms for _ in range(7))
out["hosford_fused_ms_4e6"] = ts[3]; out["hosford_plastic"] = hb.last_stats.n_plastic / n2
print(json.dumps(out))
