"""J2+Voce kernel: resident CTAs per SM (DXM_MINB 2 vs 3) across batch sizes, packed tangent."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
res = []
for n in (300_000, 1_000_000, 3_000_000, 10_000_000, 30_000_000, 100_000_000):
    for minb in (2, 3):
        code = f"""
import sys, json; sys.path.insert(0, {ROOT!r})
import dolfinx_materials_b200 as jm
m = jm.CUDAMaterial(jm.vonMisesIsotropicHardening(elasticity=jm.LinearElasticIsotropic(E=70e3, nu=0.3), yield_stress=jm.VoceHardening(sig0=350., sigu=500., b=1e3)))
n={n}; m.set_data_manager(n)
for k in range(1,4):
    m.synth_gradients(0, 1.25e-2, k, 4); m.integrate_resident(); m.data_manager.update()
m.synth_gradients(0, 1.25e-2, 4, 4)
ts=sorted(m.integrate_resident().kernel_ms for _ in range(30))
print(json.dumps(dict(n=n, minb={minb}, ms=ts[15], best=ts[0], gps=n/ts[15]*1e3, moved_gbs=472*n/ts[15]/1e6)))
"""
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, DXM_MINB=str(minb)))
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            d = dict(n=n, minb=minb, error=r.stderr[-300:])
        print(d, flush=True); res.append(d)
os.makedirs("gpurun_out", exist_ok=True); json.dump(res, open("gpurun_out/sweep_minb_n.json", "w"), indent=1)
