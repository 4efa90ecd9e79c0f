"""e2e (host arrays in / out) under N ranks on one box: where does the aggregate host link saturate, and does sending the
symmetric tangent packed (21 of 36 entries over PCIe, mirrored into the caller's (n, 36) array by host threads:
DXM_HOST_MIRROR=1) help once the AGGREGATE link -- not one rank's host memory system -- is the limit?
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 scripts/ab_e2e_scaling.py [n_per_rank]
Three hand-offs per mode, all ranks at once, max over ranks: integrate (flux + isv + Ct: 392 B/pt D2H), exchange
(flux + Ct: 336 B/pt, the QuadratureMap.update hand-off), and the packed variants of both (280 / 216 B/pt on the link)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import dolfinx_materials_b200 as jm
from dolfinx_materials_b200.material import PinnedArray

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
m = jm.CUDAMaterial(jm.vonMisesIsotropicHardening(elasticity=jm.LinearElasticIsotropic(E=70e3, nu=0.3),
                                                   yield_stress=jm.VoceHardening(sig0=350.0, sigu=500.0, b=1e3)), device=lr)
m.set_data_manager(n)
m.synth_gradients(0, 1.25e-2, 1, 1, start=rank * n)
m.integrate_resident()
g = PinnedArray((n, 6)); g.array[:] = m.device_view("strain").T.contiguous().cpu().numpy()
flux, isv, ct = PinnedArray((n, 6)), PinnedArray((n, 7)), PinnedArray((n, 36))
res = dict(world=world, n_per_rank=n)
for mirror in ("0", "1"):
    os.environ["DXM_HOST_MIRROR"] = mirror
    for name, io in (("integrate", (flux.array, isv.array, ct.array)), ("exchange", (flux.array, None, ct.array))):
        m.integrate_into(g.array, *io)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            m.integrate_into(g.array, *io)
        dt = (time.perf_counter() - t0) / 3
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = t.item()
        res[f"{name}_mirror{mirror}_ms"] = dt * 1e3
        res[f"{name}_mirror{mirror}_gps"] = world * n / dt
if rank == 0:
    print(json.dumps(res))
    os.makedirs("gpurun_out", exist_ok=True); json.dump(res, open(f"gpurun_out/ab_e2e_scaling_n{world}.json", "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
