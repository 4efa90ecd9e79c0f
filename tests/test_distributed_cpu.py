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
