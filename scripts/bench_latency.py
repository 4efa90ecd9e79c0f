"""Per-call latency of integrate() for small batches (cfg1 sizes and small meshes): host arrays in / out, the resident
call (statistics read back: a spin on the record the kernel publishes), and the asynchronous resident call."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dolfinx_materials_b200 as jm
from oracle import synth
out = []
for n in (1, 16, 1024, 10_000, 100_000, 1_000_000):
    m = jm.CUDAMaterial(jm.vonMisesIsotropicHardening(elasticity=jm.LinearElasticIsotropic(E=70e3, nu=0.3),
                                                       yield_stress=jm.VoceHardening(sig0=350.0, sigu=500.0, b=1e3)))
    m.set_data_manager(n)
    eps = synth.strain(n, 0, 1.25e-2, 1, 1)
    for _ in range(20):
        m.integrate(eps)
    reps = 500 if n <= 100_000 else 20
    t0 = time.perf_counter()
    for _ in range(reps):
        m.integrate(eps)
    dt = (time.perf_counter() - t0) / reps
    for _ in range(20):
        m.integrate_resident()
    t0 = time.perf_counter()
    for _ in range(reps):
        m.integrate_resident()
    dr = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        m.integrate_resident(wait=False)
    m.fetch_stats()
    da = (time.perf_counter() - t0) / reps
    m.enable_timing(1)
    km = min(m.integrate_resident().kernel_ms for _ in range(20))
    out.append(dict(n=n, host_call_us=dt * 1e6, resident_call_us=dr * 1e6, resident_async_us=da * 1e6, kernel_us=km * 1e3, host_gps=n / dt))
    print(out[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True); json.dump(out, open("gpurun_out/latency.json", "w"), indent=1)
