"""FeFp kernel throughput vs batch size on one B200, with SM clock / power sampled during the timed loops
(is the n-dependence a clock effect or a memory-system effect?).  Writes gpurun_out/sweep_fefp_n.json."""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dolfinx_materials_b200 as jm

res = []
for n in [int(float(a)) for a in (sys.argv[1:] or ["1e7", "2e7", "4e7", "8e7"])]:
    beh = jm.FeFpJ2Plasticity(elasticity=jm.LinearElasticIsotropic(E=70e3, nu=0.3),
                              yield_stress=jm.VoceHardening(sig0=500.0, sigu=750.0, b=1000.0))
    m = jm.CUDAMaterial(beh); m.set_data_manager(n)
    K = 4
    for k in range(1, K):
        m.synth_gradients(0, 3e-2, k, K); m.integrate_resident(); m.data_manager.update()
    m.synth_gradients(0, 3e-2, K, K)
    for _ in range(3):
        m.integrate_resident()
    smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "20"],
                           stdout=subprocess.PIPE, text=True)
    time.sleep(0.3)
    ts = []
    t0 = time.time()
    while time.time() - t0 < float(os.environ.get("DXM_SUSTAIN_S", "3.0")):
        ts.append(m.integrate_resident().kernel_ms)
    smi.terminate()
    lines = [l.split(",") for l in smi.stdout.read().strip().splitlines() if "," in l]
    clk = sorted(float(l[0]) for l in lines); pw = sorted(float(l[1]) for l in lines)
    ts.sort()
    r = dict(n=n, reps=len(ts), ms_median=ts[len(ts) // 2], ms_best=ts[0], gbs_median=976 * n / ts[len(ts) // 2] / 1e6, gbs_best=976 * n / ts[0] / 1e6,
             sm_mhz_median=clk[len(clk) // 2] if clk else None, sm_mhz_min=clk[0] if clk else None, power_w_max=pw[-1] if pw else None)
    print(r, flush=True); res.append(r)
    del m
os.makedirs("gpurun_out", exist_ok=True); json.dump(res, open("gpurun_out/sweep_fefp_n.json", "w"), indent=1)
