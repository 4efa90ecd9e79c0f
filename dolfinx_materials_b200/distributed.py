"""Multi-GPU plumbing: one process per GPU, Gauss points partitioned by contiguous cell blocks.

The constitutive update has no inter-point dependency (reference: ``jax.vmap`` over points,
``jaxmat.py:147-151``; DOLFINx already partitions cells per MPI rank, ``quadrature_map.py:66-70``), so
the data path has no collective at all.  The only exchange is the reduction of the per-call
statistics -- SUM of failed / plastic points, MAX of the local iteration count and residual -- done
with NCCL (``torch.distributed``, backend "nccl") over NVLink; ``gloo`` works for CPU-side tests.
"""

import dataclasses



_LOCAL_RANK_VARS = ("LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "MV2_COMM_WORLD_LOCAL_RANK", "MPI_LOCALRANKID",
                    "PMI_LOCAL_RANK", "SLURM_LOCALID")


def local_rank(environ=None):
    """Rank of this process on its node, from the launcher's environment (torchrun, Open MPI, MVAPICH2, Intel MPI /
    Hydra, Slurm); 0 when none of them is set.  The reference runs one dolfinx MPI rank per partition
    (``quadrature_map.py:66-70``); with one rank per GPU this is the device to use."""
    import os

    env = os.environ if environ is None else environ
    for var in _LOCAL_RANK_VARS:
        val = env.get(var)
        if val is not None and val.strip().lstrip("-").isdigit():
            return int(val)
    return 0


def default_device(device_count=None, environ=None):
    """``local_rank() % device_count`` (several ranks share a GPU when there are more ranks than GPUs, as the
    reference's GPU demo allows: ``finite_strain_elastoplasticity.py:33``)."""
    if device_count is None:
        from . import _lib

        device_count = _lib.load().dxm_device_count()
    r = local_rank(environ)
    return r % device_count if device_count and device_count > 0 else r


def shard_range(n_global, rank, world):
    """Contiguous range [start, stop) of Gauss points owned by ``rank`` (cell-major dof order,
    ``quadrature_map.py:255-260``); remainders go to the first ranks."""
    base, rem = divmod(int(n_global), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_start(n_per_rank, rank):
    """Global index of the first point of ``rank`` under weak scaling (fixed points per GPU)."""
    return int(n_per_rank) * int(rank)


def init_stats_comm(group=None, device=None):
    """Create the library's own NCCL communicator over the ranks of the (already initialised) ``torch.distributed``
    process group: rank 0 draws the unique id, ``broadcast_object_list`` carries it (works on gloo and nccl groups),
    every rank joins.  After this ``CUDAMaterial.use_global_stats()`` makes a material's per-call statistics global
    inside the update kernel's epilogue (records exchanged over peer memory) or, where peers cannot be mapped, with one
    NCCL all-gather on the material's own stream -- no host-side collective, no synchronisation.
    Returns the communicator size; 1 means nothing was created (single rank, or NCCL could not be loaded / joined on
    some rank): callers then reduce on the host (``allreduce_stats``)."""
    import ctypes

    import torch.distributed as dist

    from . import _lib

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 1
    lib = _lib.load()
    if lib.dxm_comm_size() > 1:
        return lib.dxm_comm_size()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    buf = ctypes.create_string_buffer(128)
    # every step is agreed on by all ranks: a rank that cannot load NCCL or join must not leave the others waiting in a
    # collective -- the function then returns 1 everywhere and callers fall back to a host-side reduction
    box = [None]
    if rank == 0 and lib.dxm_comm_unique_id(buf) == 0:
        box = [bytes(buf.raw)]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    if box[0] is None:
        return 1
    ident = ctypes.create_string_buffer(box[0], 128)
    joined = lib.dxm_comm_init(ident, rank, world, default_device() if device is None else int(device)) == 0
    flags = [None] * world
    dist.all_gather_object(flags, joined, group=group)
    if not all(flags):
        if joined:
            lib.dxm_comm_destroy()
        return 1
    # One node: exchange the records over peer memory inside the update kernel's epilogue instead of calling NCCL --
    # every rank maps every other rank's exchange buffer (cudaIpc over NVLink).  All ranks or none.
    import os

    if os.environ.get("DXM_STATS_P2P", "1") not in ("", "0"):
        hbuf = ctypes.create_string_buffer(64)
        ok = lib.dxm_comm_p2p_handle(hbuf) == 0
        gathered = [None] * world
        dist.all_gather_object(gathered, (ok, bytes(hbuf.raw)), group=group)
        if all(g[0] for g in gathered):
            blob = ctypes.create_string_buffer(b"".join(g[1] for g in gathered), 64 * world)
            ok = lib.dxm_comm_p2p_connect(blob) == 0
        else:
            ok = False
        flags = [None] * world
        dist.all_gather_object(flags, ok, group=group)
        if not all(flags):
            lib.dxm_comm_p2p_disable()
    return lib.dxm_comm_size()


def allreduce_stats(stats, group=None, device=None):
    """Reduce :class:`IntegrationStats` over the process group: counts are summed, iteration count
    and residual are maximised; ``kernel_ms`` becomes the max over ranks."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu"
    # one collective: gather the 6 numbers of every rank, combine locally (SUM counts, MAX the rest)
    world = dist.get_world_size(group)
    mine = torch.tensor(
        [stats.n_points, stats.n_plastic, stats.n_fail, stats.max_iter, stats.max_residual, stats.kernel_ms],
        dtype=torch.float64, device=device,
    )
    allv = torch.empty(world * 6, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(allv, mine, group=group)
    allv = allv.view(world, 6).cpu()
    s = allv[:, :3].sum(dim=0).tolist()
    m = allv[:, 3:].max(dim=0).values.tolist()
    return dataclasses.replace(
        stats, n_points=int(s[0]), n_plastic=int(s[1]), n_fail=int(s[2]), max_iter=int(m[0]), max_residual=m[1], kernel_ms=m[2]
    )


def gpu_numa_node(device):
    """NUMA node of the PCIe root the GPU hangs off (``/sys/bus/pci/devices/<bus id>/numa_node``), or None."""
    import torch

    try:
        prop = torch.cuda.get_device_properties(device)
        bus = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:  # noqa: BLE001
        return None


def bind_to_gpu_numa_node(device):
    """Pin this process to the CPUs of the NUMA node next to its GPU, so that the page-locked host buffers it
    allocates afterwards (first touch) sit behind the same PCIe root complex as the GPU.  With one process per GPU
    on a two-socket box this keeps the host<->device DMA of ``integrate`` off the inter-socket link.
    Returns ``(node, ncpus)`` or ``None`` when the topology cannot be read (nothing is changed then)."""
    import os

    node = gpu_numa_node(device)
    if node is None or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node, len(allowed)
    except Exception:  # noqa: BLE001
        return None
