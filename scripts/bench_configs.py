"""Device-resident throughput of the non-headline BASELINE.json configs on one B200 (cfg1, cfg3, cfg4)
plus cfg2 variants (sorted vs shuffled plastic points).  Writes gpurun_out/configs.json.
These are parity-test workloads, not bench.py lines; numbers are kernel times from CUDA events."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dolfinx_materials_b200 as jm

def timeit(m, reps=8):
    ts = []
    for _ in range(reps):
        s = m.integrate_resident(); ts.append(s.kernel_ms)
    ts = sorted(ts[2:]); return ts[len(ts) // 2], s

out = []
el = jm.LinearElasticIsotropic(E=70e3, nu=0.3)

# cfg1: uniaxial tension analogue, J2 linear hardening, all points on the same path, n = 1e6
n = 1_000_000
m = jm.CUDAMaterial(jm.vonMisesIsotropicHardening(elasticity=el, yield_stress=jm.LinearHardening(sig0=250.0, H=1e-6)))
m.set_data_manager(n)
import torch
g = m.gradient_buffer()
for exx in np.linspace(0, 2e-2, 11)[1:]:
    g = m.gradient_buffer(); g.zero_(); g[0].fill_(exx); g[1].fill_(-0.45 * exx); torch.cuda.synchronize()
    ms, s = timeit(m, 4); m.data_manager.update()
out.append(dict(cfg="cfg1 J2-linear uniaxial path", n=n, ms=ms, gps=n / ms * 1e3, gbs=592 * n / ms / 1e6, plastic=s.n_plastic / n))

# cfg3: FeFp + Voce, n = 1e7, random F = I + sG
n = 10_000_000
m = jm.CUDAMaterial(jm.FeFpJ2Plasticity(elasticity=el, yield_stress=jm.VoceHardening(sig0=500.0, sigu=750.0, b=1000.0)))
m.set_data_manager(n)
for k in range(1, 4):
    m.synth_gradients(0, 3e-2, k, 4); m.integrate_resident(); m.data_manager.update()
m.synth_gradients(0, 3e-2, 4, 4)
ms, s = timeit(m)
out.append(dict(cfg="cfg3 FeFp+Voce random", n=n, ms=ms, gps=n / ms * 1e3, gbs=976 * n / ms / 1e6, plastic=s.n_plastic / n, max_iter=s.max_iter))
del m

# cfg4: heterogeneous batch, per-point properties, block-contiguous classes, n = 1e7
cls = np.zeros(n, dtype=np.int8); cls[int(0.6 * n):int(0.9 * n)] = 1; cls[int(0.9 * n):] = 2
props = {"E": np.where(cls == 1, 90e3, 70e3), "nu": np.where(cls == 1, 0.25, 0.3), "sig0": np.where(cls == 2, np.inf, 200.0),
         "H": np.where(cls == 0, 10.0, 0.0), "sigu": np.where(cls == 1, 300.0, np.where(cls == 2, np.inf, 200.0)),
         "b": np.where(cls == 1, 10.0, 0.0)}
m = jm.CUDAMaterial(jm.vonMisesIsotropicHardening(elasticity=el, yield_stress=jm.VoceHardening(sig0=200.0, sigu=200.0, b=0.0)))
m.set_data_manager(n)
for k, v in props.items():
    m.update_material_property(k, v)
for k in range(1, 4):
    m.synth_gradients(0, 1.25e-2, k, 4); m.integrate_resident(); m.data_manager.update()
m.synth_gradients(0, 1.25e-2, 4, 4)
ms, s = timeit(m)
out.append(dict(cfg="cfg4 per-point properties (60% J2-linear, 30% Voce, 10% elastic)", n=n, ms=ms, gps=n / ms * 1e3,
                gbs=(592 + 48) * n / ms / 1e6, note="640 B/pt incl. 6 property reads", plastic=s.n_plastic / n, max_iter=s.max_iter))
del m

# cfg4 faithful variant: two handles on one GPU launched back to back (multimaterials.py:265-273)
na, nb = int(0.7 * n), n - int(0.7 * n)
ma = jm.CUDAMaterial(jm.vonMisesIsotropicHardening(elasticity=el, yield_stress=jm.LinearHardening(sig0=200.0, H=10.0)))
mb = jm.CUDAMaterial(jm.vonMisesIsotropicHardening(elasticity=jm.LinearElasticIsotropic(E=90e3, nu=0.25),
                                                   yield_stress=jm.VoceHardening(sig0=200.0, sigu=300.0, b=10.0)))
ma.set_data_manager(na); mb.set_data_manager(nb)
for mm, st in ((ma, 0), (mb, na)):
    mm.synth_gradients(0, 1.25e-2, 4, 4, start=st)
tot = []
for _ in range(6):
    a = ma.integrate_resident(); b = mb.integrate_resident(); tot.append(a.kernel_ms + b.kernel_ms)
ms = sorted(tot[2:])[2]
out.append(dict(cfg="cfg4 two handles back-to-back (matrix J2-linear + inclusions Voce)", n=n, ms=ms, gps=n / ms * 1e3, gbs=592 * n / ms / 1e6))
del ma, mb

# cfg4 with the demo's own matrix law: Hosford (a = 10) + linear hardening in the matrix handle (70 % of the points,
# IsotropicPlasticHosfordFlowLinear.mfront), J2 + Voce in the inclusions handle -- two handles back to back
mh = jm.CUDAMaterial(jm.GeneralIsotropicHardening(elasticity=el, yield_stress=jm.LinearHardening(sig0=200.0, H=10.0),
                                                  equivalent_stress=jm.Hosford(a=10)))
mb = jm.CUDAMaterial(jm.vonMisesIsotropicHardening(elasticity=jm.LinearElasticIsotropic(E=90e3, nu=0.25),
                                                   yield_stress=jm.VoceHardening(sig0=200.0, sigu=300.0, b=10.0)))
mh.set_data_manager(na); mb.set_data_manager(nb)
for mm, st in ((mh, 0), (mb, na)):
    for k in range(1, 4):
        mm.synth_gradients(0, 1.25e-2, k, 4, start=st); mm.integrate_resident(); mm.data_manager.update()
    mm.synth_gradients(0, 1.25e-2, 4, 4, start=st)
tot, th = [], []
for _ in range(6):
    a = mh.integrate_resident(); b = mb.integrate_resident(); tot.append(a.kernel_ms + b.kernel_ms); th.append(a.kernel_ms)
ms = sorted(tot[2:])[2]; msh = sorted(th[2:])[2]
out.append(dict(cfg="cfg4 faithful: matrix Hosford(a=10)+linear handle, inclusions J2+Voce handle, back to back", n=n, ms=ms,
                gps=n / ms * 1e3, hosford_n=na, hosford_ms=msh, hosford_gps=na / msh * 1e3, hosford_gbs_moved=472 * na / msh / 1e6,
                hosford_plastic=a.n_plastic / na, hosford_max_iter=a.max_iter, hosford_fail=a.n_fail))
del mh, mb
# Hosford kernel alone over the plastic fraction (amplitude sweep), n = 1e7: fused kernel vs tiled kernel (stream a 1024-point
# tile + CTA-local candidate queue + packed local solves; r01g/r01h also ran a device-wide queue + second kernel: slower), (r01g also ran the local-solve kernel at 4 resident CTAs per SM / 128 registers: 0-20 % slower)
mh = jm.CUDAMaterial(jm.GeneralIsotropicHardening(elasticity=el, yield_stress=jm.LinearHardening(sig0=200.0, H=10.0)))
mh.set_data_manager(n)
for split, minb in (("0", "4"), ("1", "3")):  # register targets: the per-kernel defaults (dxm_hosford_api.cu)
    os.environ["DXM_HOS_SPLIT"] = split
    for amp in (2e-3, 4e-3, 8e-3, 1.25e-2, 5e-2):
        mh.data_manager.revert(); mh.synth_gradients(0, amp, 1, 1)
        ms, s = timeit(mh)
        out.append(dict(cfg=f"Hosford a=10 alone, virgin state, amp {amp}", split=int(split), minb=int(minb), n=n, ms=ms, gps=n / ms * 1e3,
                        gbs_moved=472 * n / ms / 1e6, plastic=s.n_plastic / n, max_iter=s.max_iter, fail=s.n_fail))
del os.environ["DXM_HOS_SPLIT"]
del mh

# the same batch with the plastic points contiguous (sorted by strain amplitude, i.e. as a mesh with a
# localised plastic zone presents them) vs interleaved at random (the synthetic default) -- the difference is what lane
# divergence costs each kernel.  One increment from the virgin state, so that permuting the gradients permutes nothing else.
import torch
n2 = 10_000_000
for name, beh, amp in (("J2+Voce", jm.vonMisesIsotropicHardening(elasticity=el, yield_stress=jm.VoceHardening(sig0=350.0, sigu=500.0, b=1e3)), 6e-3),
                       ("Hosford a=10 (auto fused/tiled)", jm.GeneralIsotropicHardening(elasticity=el, yield_stress=jm.LinearHardening(sig0=200.0, H=10.0)), 4e-3),
                       ("FeFp+Voce", jm.FeFpJ2Plasticity(elasticity=el, yield_stress=jm.VoceHardening(sig0=500.0, sigu=750.0, b=1000.0)), 1.2e-2)):
    mm = jm.CUDAMaterial(beh); mm.set_data_manager(n2)
    mm.synth_gradients(0, amp, 1, 1)
    g = mm.gradient_buffer()
    dev = g.clone()
    if dev.shape[0] == 9: dev[:3] -= 1.0
    perm = torch.argsort((dev * dev).sum(dim=0))
    del dev
    res = {}
    for order in ("shuffled", "sorted"):
        if order == "sorted":
            g.copy_(g[:, perm].clone()); torch.cuda.synchronize()
        mm.integrate_resident()  # auto modes key on the previous call's plastic fraction
        ms, s = timeit(mm)
        res[order] = ms
    out.append(dict(cfg=f"divergence exposure: {name}, n=1e7, one increment from the virgin state", shuffled_ms=res["shuffled"],
                    sorted_ms=res["sorted"], plastic=s.n_plastic / n2, sorted_speedup=res["shuffled"] / res["sorted"]))
    del mm, g, perm
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True); json.dump(out, open("gpurun_out/configs.json", "w"), indent=1)
