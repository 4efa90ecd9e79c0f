"""Platform probe for the e2e leg at N GPUs (run under torchrun): aggregate pinned-memory D2H / H2D bandwidth with
all ranks copying at once, plus the host topology the ranks see.  Separates what the box can move from what the
dxm_integrate pipeline achieves.  Rank 0 writes gpurun_out/probe_pcie_n<N>.json."""
import json, os, subprocess, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
from dolfinx_materials_b200.distributed import bind_to_gpu_numa_node, gpu_numa_node
bound = None if os.environ.get("NO_BIND") else bind_to_gpu_numa_node(local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nbytes = 1 << 30
h = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); h.fill_(1)
d = torch.empty(nbytes, dtype=torch.uint8, device="cuda"); d.fill_(2)
s2 = torch.cuda.Stream()


def run(kind, reps=8):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        if kind in ("d2h", "both"):
            h.copy_(d, non_blocking=True)
        if kind in ("h2d", "both"):
            with torch.cuda.stream(s2):
                d2.copy_(h2, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n_dir = 2 if kind == "both" else 1
    return world * reps * nbytes * n_dir / t.item() / 1e9


h2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); d2 = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
out = {}
for kind in ("d2h", "h2d", "both"):
    run(kind, 2)
    out[kind + "_aggregate_gbs"] = run(kind)
info = dict(rank=rank, numa_node_of_gpu=gpu_numa_node(local), bound=bound, affinity=len(os.sched_getaffinity(0)))
gathered = [None] * world
if world > 1:
    dist.all_gather_object(gathered, info)
else:
    gathered = [info]
if rank == 0:
    out["world"] = world
    out["ranks"] = gathered
    for cmd in (["nvidia-smi", "topo", "-m"], ["lscpu"], ["cat", "/proc/meminfo"]):
        try:
            txt = subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout
            out[" ".join(cmd)] = txt if cmd[0] != "cat" else "\n".join(txt.splitlines()[:3])
        except Exception as e:  # noqa: BLE001
            out[" ".join(cmd)] = repr(e)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/probe_pcie_n{world}.json", "w"), indent=1)
    print({k: v for k, v in out.items() if k.endswith("gbs")}, gathered[:2])
if world > 1:
    dist.destroy_process_group()
