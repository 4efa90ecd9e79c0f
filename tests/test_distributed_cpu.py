"""N > 1 host logic on CPU: world_size-2 gloo process group (shard ranges, statistics all-reduce)."""
import os
import sys

import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from dolfinx_materials_b200.distributed import allreduce_stats, shard_range, shard_start
    from dolfinx_materials_b200.material import IntegrationStats

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = shard_range(1001, rank, world)
    s = IntegrationStats(n_points=hi - lo, n_plastic=10 * (rank + 1), n_fail=rank, max_iter=3 + rank,
                         max_residual=1e-10 * (rank + 1), kernel_ms=1.0 + rank)
    r = allreduce_stats(s)
    q.put((rank, lo, hi, r, shard_start(500, rank)))
    dist.destroy_process_group()


def test_shard_and_allreduce_world2():
    world, port = 2, 29611
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, s0, st0), (r1, lo1, hi1, s1, st1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 501, 501, 1001)  # contiguous, remainder to the first rank
    assert (st0, st1) == (0, 500)
    for s in (s0, s1):
        assert s.n_points == 1001 and s.n_plastic == 30 and s.n_fail == 1
        assert s.max_iter == 4 and s.max_residual == 2e-10 and s.kernel_ms == 2.0


def test_shard_range_covers_everything():
    from dolfinx_materials_b200.distributed import shard_range

    for n in (1, 7, 8, 100003):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_numa_binding_is_best_effort_without_a_gpu():
    """No CUDA device / no exposed topology: the helpers report None and leave the affinity mask alone."""
    import os

    from dolfinx_materials_b200.distributed import bind_to_gpu_numa_node, gpu_numa_node

    before = os.sched_getaffinity(0)
    assert gpu_numa_node(0) is None
    assert bind_to_gpu_numa_node(0) is None
    assert os.sched_getaffinity(0) == before


def _asm_worker(rank, world, port, q):
    """Rank-sharded assembly, host logic: each rank assembles its contiguous block of cells with the constrained rows
    deferred, values / rhs are summed over the group, constraints are applied once (what
    AssembledSystem.assemble_sharded does with the CUDA kernels + NCCL)."""
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import torch.distributed as dist

    from dolfinx_materials_b200.distributed import shard_range
    from oracle import fe_forms as ff
    from oracle import fe_gradient as fg

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    coords, gd, ud, nodes = fg.box_tets(3, 2, 2, 2)
    dphi, w = fg.tet_dphi(fg.TET_QP_DEG2, 2), np.full(4, 1.0 / 24.0)
    nc = len(gd)
    rng = np.random.default_rng(11)  # same data on every rank
    flux = rng.standard_normal((nc * 4, 9))
    ct = rng.standard_normal((nc * 4, 81))
    n = 3 * len(nodes)
    bc = np.repeat(nodes[:, 0] < 0.05, 3)
    lift = np.where(bc, 1e-3 * np.cos(np.arange(n)), 0.0)
    c0, c1 = shard_range(nc, rank, world)
    fe, ke = ff.element_forms(coords, gd[c0:c1], ud[c0:c1], dphi, w, flux[4 * c0:4 * c1], ct[4 * c0:4 * c1], 1, 3)
    b, A = ff.assemble(ud[c0:c1], fe, ke, len(nodes), 3, bc=bc, lift=lift, constrain=False)
    dense = torch.from_numpy(A.toarray())
    bt = torch.from_numpy(b)
    dist.all_reduce(dense)
    dist.all_reduce(bt)
    import scipy.sparse as sp

    A2, b2 = ff.apply_constraints(sp.csr_matrix(dense.numpy()), bt.numpy(), bc, lift)
    fe_all, ke_all = ff.element_forms(coords, gd, ud, dphi, w, flux, ct, 1, 3)
    b_ref, A_ref = ff.assemble(ud, fe_all, ke_all, len(nodes), 3, bc=bc, lift=lift)
    scale = np.abs(ke_all).max()
    q.put((rank, c0, c1, float(abs(A2 - A_ref).max() / scale), float(np.abs(b2 - b_ref).max() / np.abs(b_ref).max())))
    dist.destroy_process_group()


def test_rank_sharded_assembly_world2():
    world, port = 2, 29613
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_asm_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 72
    for _, _, _, ea, eb in res:
        assert ea < 1e-13 and eb < 1e-13


def test_local_rank_and_default_device_from_launcher_environment():
    from dolfinx_materials_b200.distributed import default_device, local_rank

    assert local_rank({}) == 0
    assert local_rank({"OMPI_COMM_WORLD_LOCAL_RANK": "3"}) == 3
    assert local_rank({"LOCAL_RANK": "5", "SLURM_LOCALID": "1"}) == 5  # torchrun wins over the scheduler
    assert local_rank({"SLURM_LOCALID": "x"}) == 0
    assert default_device(8, {"LOCAL_RANK": "11"}) == 3  # more ranks than GPUs: ranks share devices
    assert default_device(0, {"LOCAL_RANK": "2"}) == 2  # no device visible: dxm_create will say so


def test_device_count_without_a_gpu():
    import torch

    from dolfinx_materials_b200 import _lib, build

    build.build_library()
    n = _lib.load().dxm_device_count()
    assert n == (torch.cuda.device_count() if torch.cuda.is_available() else 0)
