"""Re-tune the launch shape of the J2+Voce kernel after the packed-tangent change (25 read + 34 write streams):
tiles per CTA (DXM_TPB), resident CTAs the register allocation targets (DXM_MINB), access width (DXM_PPT)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = f"""
import sys, json; sys.path.insert(0, {ROOT!r})
import dolfinx_materials_b200 as jm
m = jm.CUDAMaterial(jm.vonMisesIsotropicHardening(elasticity=jm.LinearElasticIsotropic(E=70e3, nu=0.3), yield_stress=jm.VoceHardening(sig0=350., sigu=500., b=1e3)))
n=100_000_000; m.set_data_manager(n)
for k in range(1,4):
    m.synth_gradients(0, 1.25e-2, k, 4); m.integrate_resident(); m.data_manager.update()
m.synth_gradients(0, 1.25e-2, 4, 4)
ts=sorted(m.integrate_resident().kernel_ms for _ in range(10))
print(json.dumps(dict(n=n, ms=ts[4], best=ts[0], gps=n/ts[4]*1e3, moved_gbs=472*n/ts[4]/1e6)))
"""
res = []
ENVS = ([dict(DXM_TPB=str(t)) for t in (1, 2, 3, 4, 6, 8, 16)] + [dict(DXM_MINB=str(b)) for b in (1, 2, 4)]
        + [dict(DXM_PPT="2"), dict(DXM_GRID="3"), dict(DXM_GRID="6"), dict(DXM_MINB="2", DXM_TPB="16")])
for env in ENVS:
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, **env))
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        d = dict(error=r.stderr[-300:])
    d["env"] = env
    print(d, flush=True); res.append(d)
os.makedirs("gpurun_out", exist_ok=True); json.dump(res, open("gpurun_out/sweep_tpb_packed.json", "w"), indent=1)
