"""Hosford kernels A/B on one B200: resident CTAs per SM the register allocation targets (DXM_HOS_MINB = 3: 168
registers, 4: 128) x fused / tiled launch (DXM_HOS_SPLIT) over the plastic fraction (amplitude sweep), n = 1e7, a = 10."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dolfinx_materials_b200 as jm

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
mh = jm.CUDAMaterial(jm.GeneralIsotropicHardening(elasticity=jm.LinearElasticIsotropic(E=70e3, nu=0.3),
                                                  yield_stress=jm.LinearHardening(sig0=200.0, H=10.0)))
mh.set_data_manager(n)
out = []
for minb in ("3", "4"):
    for split in ("0", "1"):
        os.environ["DXM_HOS_MINB"], os.environ["DXM_HOS_SPLIT"] = minb, split
        for amp in (2e-3, 2.5e-3, 3e-3, 4e-3, 8e-3, 1.25e-2, 5e-2):
            mh.data_manager.revert(); mh.synth_gradients(0, amp, 1, 1)
            ts = []
            for _ in range(7):
                s = mh.integrate_resident(); ts.append(s.kernel_ms)
            ms = sorted(ts[2:])[2]
            out.append(dict(minb=int(minb), tiled=int(split), amp=amp, ms=ms, gps=n / ms * 1e3, plastic=s.n_plastic / n,
                            max_iter=s.max_iter, fail=s.n_fail))
            print(out[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True); json.dump(out, open("gpurun_out/ab_hosford.json", "w"), indent=1)
