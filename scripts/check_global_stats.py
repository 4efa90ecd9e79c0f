"""Multi-GPU check of the in-stream statistics reduction (run under torchrun, one rank per GPU):
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/check_global_stats.py
Every rank integrates its own shard (different sizes, one rank gets a failing point); with global statistics on, every
rank must read the same record = SUM / MAX of the ranks' local records, and the per-call wall time of the resident call
(kernel + NCCL all-gather of 64 B per rank + one-warp fold, all on the handle's stream) is reported next to the local one."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import dolfinx_materials_b200 as jm
from dolfinx_materials_b200 import distributed as dd
from oracle import synth

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
assert dd.init_stats_comm() == world
p2p = bool(jm._lib.load().dxm_comm_p2p_enabled())

def material(n):
    m = jm.CUDAMaterial(jm.vonMisesIsotropicHardening(elasticity=jm.LinearElasticIsotropic(E=70e3, nu=0.3),
                                                       yield_stress=jm.VoceHardening(sig0=350.0, sigu=500.0, b=1e3)))
    m.set_data_manager(n)
    return m

n = 10_000 + 1000 * rank
m = material(n)
eps = synth.strain(n, rank, 1.25e-2, 1, 1)
if rank == world - 1:
    eps[7, 0] = np.nan
import warnings
warnings.simplefilter("ignore")
m.integrate(eps)
loc = m.last_stats
gathered = [None] * world
dist.all_gather_object(gathered, (loc.n_points, loc.n_plastic, loc.n_fail, loc.max_iter))
m.use_global_stats()
for path in ("host", "resident", "async"):
    if path == "host":
        m.integrate(eps); g = m.last_stats
    elif path == "resident":
        g = m.integrate_resident()
    else:
        m.integrate_resident(wait=False); g = m.fetch_stats()
    assert g.n_points == sum(x[0] for x in gathered), (path, g, gathered)
    assert g.n_plastic == sum(x[1] for x in gathered)
    assert g.n_fail == sum(x[2] for x in gathered) == 1
    assert g.max_iter == max(x[3] for x in gathered)
# latency of the resident call with and without the collective (clean data)
res = {}
for n in (528_000, 10_000):
    m2 = material(n); m2.synth_gradients(rank, 1.25e-2, 1, 1)
    for glob in (0, 1):
        m2.use_global_stats(glob)
        for _ in range(50):
            m2.integrate_resident()
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(300):
            m2.integrate_resident()
        res[f"n{n}_{'global' if glob else 'local'}_us"] = (time.perf_counter() - t0) / 300 * 1e6
    m2.enable_timing(1); m2.use_global_stats(0)
    res[f"n{n}_kernel_us"] = min(m2.integrate_resident().kernel_ms for _ in range(20)) * 1e3
    del m2
if rank == 0:
    out = dict(world=world, ok=True, exchange="peer memory (in the kernel epilogue)" if p2p else "NCCL all-gather + publish kernel", **res)
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/global_stats_n{world}_{'p2p' if p2p else 'nccl'}.json", "w"), indent=1)
dist.barrier()
lib = jm._lib.load()
del m
lib.dxm_comm_destroy()
dist.destroy_process_group()
