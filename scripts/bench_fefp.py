"""FeFp (cfg3) device-resident throughput on one B200: n points, random F = I + s G histories."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dolfinx_materials_b200 as jm

def hbm_peak():
    """the pod's measured HBM copy bandwidth (MEASURED_PEAKS.json, driver-written), GB/s"""
    try:
        return float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        return 6650.0  # B200_PROFILING.md fallback


n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
amp = float(sys.argv[2]) if len(sys.argv) > 2 else 3e-2
beh = jm.FeFpJ2Plasticity(elasticity=jm.LinearElasticIsotropic(E=70e3, nu=0.3),
                          yield_stress=jm.VoceHardening(sig0=500.0, sigu=750.0, b=1000.0))
m = jm.CUDAMaterial(beh); m.set_data_manager(n)
K = 4
for k in range(1, K):
    m.synth_gradients(0, amp, k, K); m.integrate_resident(); m.data_manager.update()
m.synth_gradients(0, amp, K, K)
ts = []
for i in range(10):
    s = m.integrate_resident(); ts.append(s.kernel_ms)
ts = sorted(ts[2:]); ms = ts[len(ts) // 2]
print(json.dumps(dict(kind="fefp", n=n, amp=amp, ms=ms, gps=n / ms * 1e3, gbs=976 * n / ms / 1e6, frac_hbm=976 * n / ms / 1e6 / hbm_peak(), hbm_peak_gbs=hbm_peak(),
                      plastic=s.n_plastic / n, max_iter=s.max_iter, n_fail=s.n_fail, max_resid=s.max_residual)))
