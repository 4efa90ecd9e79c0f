"""A/B of the packed-tangent hand-off on the host path: e2e `CUDAMaterial.integrate(host gradients)` points/s with
DXM_HOST_MIRROR=0 (device expands the symmetric tangent, 36 doubles/point over PCIe) and =1 (21 doubles/point over
PCIe, host threads mirror).  The pool size is fixed per process (DXM_HOST_THREADS), so each thread count is its own
process:  python scripts/ab_host_mirror.py [threads ...]  -> gpurun_out/ab_host_mirror.json"""
import json, os, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import numpy as np
    import dolfinx_materials_b200 as jm
    from dolfinx_materials_b200.material import PinnedArray
    from oracle import synth

    res = []
    for n in (100_000, 1_000_000, 10_000_000):
        beh = jm.vonMisesIsotropicHardening(elasticity=jm.LinearElasticIsotropic(E=70e3, nu=0.3),
                                            yield_stress=jm.VoceHardening(sig0=350.0, sigu=500.0, b=1e3))
        g = PinnedArray((n, 6))
        g.array[...] = synth.strain(n, 0, 1.25e-2, 1, 1)
        ref = None
        for mirror in ("0", "1"):
            os.environ["DXM_HOST_MIRROR"] = mirror
            m = jm.CUDAMaterial(beh)
            m.set_data_manager(n)
            for _ in range(3):
                flux, isv, ct = m.integrate(g.array)
            reps = 5 if n >= 10_000_000 else 20
            t0 = time.perf_counter()
            for _ in range(reps):
                flux, isv, ct = m.integrate(g.array)
            dt = (time.perf_counter() - t0) / reps
            if ref is None:
                ref = ct.copy()
            same = bool(np.array_equal(ref, ct))
            res.append(dict(n=n, mirror=int(mirror), threads=os.environ.get("DXM_HOST_THREADS", "default"),
                            ms=dt * 1e3, gps=n / dt, identical=same))
            print(res[-1], flush=True)
            del m
    print("RESULT " + json.dumps(res), flush=True)


if __name__ == "__main__":
    if os.environ.get("DXM_AB_CHILD"):
        child()
        sys.exit(0)
    out = []
    for th in (sys.argv[1:] or ["default"]):
        env = dict(os.environ, DXM_AB_CHILD="1")
        if th != "default":
            env["DXM_HOST_THREADS"] = th
        r = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True)
        sys.stderr.write(r.stderr[-2000:])
        for line in r.stdout.splitlines():
            if line.startswith("RESULT "):
                out += json.loads(line[7:])
            else:
                print(line, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ab_host_mirror.json"), "w"), indent=1)
