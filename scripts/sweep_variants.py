"""Kernel-variant sweep on one B200 (run through gpurun): PPT x MINB for the J2+Voce kernel, plus the
copy / FP64 peaks measured by the library's own microbenchmarks.  Writes gpurun_out/sweep.json."""
import ctypes, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def one(ppt, minb, n, kind):
    code = f"""
import os, sys, json
os.environ['DXM_PPT']='{ppt}'; os.environ['DXM_MINB']='{minb}'
sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
import dolfinx_materials_b200 as jm
el = jm.LinearElasticIsotropic(E=70e3, nu=0.3)
kind={kind!r}
if kind=='voce': beh = jm.vonMisesIsotropicHardening(elasticity=el, yield_stress=jm.VoceHardening(sig0=350., sigu=500., b=1e3))
elif kind=='linear': beh = jm.vonMisesIsotropicHardening(elasticity=el, yield_stress=jm.LinearHardening(sig0=350., H=5e3))
else: beh = jm.ElasticBehavior(elasticity=el)
m = jm.CUDAMaterial(beh); n={n}; m.set_data_manager(n)
K=4
for k in range(1,K):
    m.synth_gradients(0, 1.25e-2, k, K); m.integrate_resident(); m.data_manager.update()
m.synth_gradients(0, 1.25e-2, K, K)
ts=[]
for i in range(8):
    s = m.integrate_resident(); ts.append(s.kernel_ms)
ts=sorted(ts[2:])
ms = ts[len(ts)//2]
print(json.dumps(dict(kind=kind, ppt={ppt}, minb={minb}, n=n, ms=ms, best_ms=ts[0], gps=n/ms*1e3, gbs=592*n/ms/1e6, plastic=s.n_plastic/n, max_iter=s.max_iter)))
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    if r.returncode != 0:
        return dict(error=r.stderr[-500:], ppt=ppt, minb=minb)
    return json.loads(r.stdout.strip().splitlines()[-1])

if __name__ == "__main__":
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
    from dolfinx_materials_b200 import _lib
    lib = _lib.load()
    out = {}
    v = ctypes.c_double()
    _lib.check(lib.dxm_copy_peak(0, 1 << 31, ctypes.byref(v))); out["copy_gbs"] = v.value
    _lib.check(lib.dxm_fp64_peak(0, ctypes.byref(v))); out["fp64_tflops"] = v.value
    print(out, flush=True)
    res = []
    for kind in ["voce", "linear", "elastic"]:
        for ppt, minb in [(1, 1), (1, 2), (1, 3), (1, 4), (2, 1), (2, 2), (2, 3)]:
            if kind != "voce" and minb != 2:
                continue
            r = one(ppt, minb, n, kind); print(r, flush=True); res.append(r)
    out["variants"] = res
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/sweep.json", "w"), indent=1)
