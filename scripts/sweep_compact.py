"""A/B of the block-level compaction of the plastic points' local Newton solves (DXM_COMPACT=0|1) on one B200,
J2+Voce (n = 5e7) and FeFp (n = 2e7), over synthetic-history amplitudes that span the plastic fraction.
Writes gpurun_out/sweep_compact.json."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def one(kind, compact, n, amp):
    code = f"""
import os, sys, json
os.environ['DXM_COMPACT']='{compact}'
sys.path.insert(0, {ROOT!r})
import dolfinx_materials_b200 as jm
el = jm.LinearElasticIsotropic(E=70e3, nu=0.3)
kind={kind!r}; n={n}; amp={amp}
if kind=='voce':
    beh = jm.vonMisesIsotropicHardening(elasticity=el, yield_stress=jm.VoceHardening(sig0=350., sigu=500., b=1e3)); B=592
else:
    beh = jm.FeFpJ2Plasticity(elasticity=el, yield_stress=jm.VoceHardening(sig0=500., sigu=750., b=1000.)); B=976
m = jm.CUDAMaterial(beh); m.set_data_manager(n)
K=4
for k in range(1,K):
    m.synth_gradients(0, amp, k, K); m.integrate_resident(); m.data_manager.update()
m.synth_gradients(0, amp, K, K)
ts=[]
for i in range(10):
    s = m.integrate_resident(); ts.append(s.kernel_ms)
ts=sorted(ts[2:]); ms = ts[len(ts)//2]
print(json.dumps(dict(kind=kind, compact={compact}, n=n, amp=amp, ms=ms, gps=n/ms*1e3, gbs=B*n/ms/1e6, plastic=s.n_plastic/n, max_iter=s.max_iter, n_fail=s.n_fail)))
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    if r.returncode != 0:
        return dict(error=r.stderr[-500:], kind=kind, compact=compact, amp=amp)
    return json.loads(r.stdout.strip().splitlines()[-1])


if __name__ == "__main__":
    res = []
    for kind, n, amps in (("voce", 50_000_000, (4e-3, 6e-3, 1.25e-2, 3e-2)), ("fefp", 20_000_000, (6e-3, 1e-2, 1.6e-2, 3e-2))):
        for amp in amps:
            for compact in (0, 1):
                r = one(kind, compact, n, amp); print(r, flush=True); res.append(r)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/sweep_compact.json", "w"), indent=1)
