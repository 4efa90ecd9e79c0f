"""Grid-size sweep (DXM_GRID = CTAs per SM, 0 = one CTA per tile) for the J2+Voce and FeFp kernels."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
res = []
GRIDS = [int(g) for g in os.environ.get("GRIDS", "2,4,8,16,64,0").split(",")]
KINDS = os.environ.get("KINDS", "voce,fefp").split(",")
for kind, n in (("voce", 100_000_000), ("fefp", 40_000_000)):
    if kind not in KINDS:
        continue
    for grid in GRIDS:
        env = dict(os.environ, DXM_GRID=str(grid)) if grid >= 0 else dict(os.environ, DXM_TPB=str(-grid))
        if kind == "voce":
            code = f"""
import sys, json; sys.path.insert(0, {ROOT!r})
import dolfinx_materials_b200 as jm
m = jm.CUDAMaterial(jm.vonMisesIsotropicHardening(elasticity=jm.LinearElasticIsotropic(E=70e3, nu=0.3), yield_stress=jm.VoceHardening(sig0=350., sigu=500., b=1e3)))
n={n}; m.set_data_manager(n)
for k in range(1,4):
    m.synth_gradients(0, 1.25e-2, k, 4); m.integrate_resident(); m.data_manager.update()
m.synth_gradients(0, 1.25e-2, 4, 4)
ts=sorted(m.integrate_resident().kernel_ms for _ in range(8))
print(json.dumps(dict(kind='voce', grid={grid}, n=n, ms=ts[3], gps=n/ts[3]*1e3, gbs=592*n/ts[3]/1e6)))
"""
            r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
        else:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts/bench_fefp.py"), str(n), "3e-2"], capture_output=True, text=True, env=env)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1]); d["grid"] = grid
        except Exception:
            d = dict(kind=kind, grid=grid, error=r.stderr[-300:])
        print(d, flush=True); res.append(d)
os.makedirs("gpurun_out", exist_ok=True); json.dump(res, open("gpurun_out/sweep_grid.json", "w"), indent=1)
