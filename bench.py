#!/usr/bin/env python
"""bench.py -- Gauss-point updates/s of the batched constitutive update (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: numpy oracle on all host cores

Workload (cfg2 of BASELINE.json / SURVEY.md 8(d)): 3-D small-strain J2 plasticity with Voce hardening
(E=70e3, nu=0.3, sig0=350, sigu=500, b=1e3 -- plane_elastoplasticity.py:60-69), fp64, 1e8 synthetic Gauss
points PER GPU (weak scaling; contiguous point ranges per rank), proportional strain histories from the
counter-based recipe of oracle/synth.py (amp 1.25e-2, 4 increments, ~64 % plastic points, <= 5 local
Newton iterations).  One step = one `integrate` over all points at the last increment, starting from
the state reached after the first three (exactly what each global Newton iteration does:
quadrature_map.py:320-321), including the device-side statistics reduction and, for N > 1, the NCCL
all-reduce of the failure / active-set counts and residual maximum.

`value`  : device-resident throughput (inputs and outputs stay in HBM, SoA).
`e2e`    : same metric through the reference-facing call `CUDAMaterial.integrate(host gradients)` ->
           host (flux, isv, Ct): pinned host buffers, H2D + D2H inside the timed region.
`roofline`: algorithmic bytes (592 B / Gauss point: the full 36-entry tangent of the reference boundary) /
           average kernel time measured with CUDA events on the launching stream inside the timed region, against
           MEASURED_PEAKS.json hbm_gbs.  The kernel stores each unique entry of the symmetric tangent once
           (472 B / point of DRAM traffic, `traffic`), so `frac` can exceed 1; `moved_frac` is the fraction of the
           HBM peak the bytes actually moved account for.
`cpu_baseline`: the numpy oracle timed on the box's host cores on a bounded sample of the same workload
           (`c_port`: the plain-C oracle on the same cores, for scale).
"""

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PROPS = dict(E=70e3, nu=0.3, sig0=350.0, sigu=500.0, b=1e3)
AMP, KINC, SEED = 1.25e-2, 4, 0
BYTES_PER_GP = 592  # algorithmic: 25 doubles read + 49 written
BYTES_MOVED_PER_GP = 472  # what the kernel moves: the symmetric tangent is stored once (21 of its 36 entries)
METRIC = "GP updates/s (fp64 J2 return map + Ct)"
UNIT = "GP/s"


# ------------------------------------------------------------------------------------------------
# CPU arm / baseline: numpy oracle, one process per host core
# ------------------------------------------------------------------------------------------------
def _cpu_setup(args):
    """state after increments 1..K-1 for points [start, start+n) -- untimed"""
    from oracle import small_strain as ss
    from oracle import synth

    start, n = args
    st = ss.zero_state(n)
    t0 = time.perf_counter()
    for k in range(1, KINC):
        out = ss.integrate(synth.strain(n, SEED, AMP, k, KINC, start=start), st, PROPS)
        st = ss.advance(out)
    rate = n * (KINC - 1) / (time.perf_counter() - t0)
    return st, rate


def _cpu_worker(conn):
    """Owns a fixed set of point chunks; builds their state once (untimed), then replays the last
    increment on request."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import small_strain as ss
    from oracle import synth

    states = {}
    while True:
        msg = conn.recv()
        if msg[0] == "stop":
            break
        if msg[0] == "prepare":
            rates = []
            states = {t: states[t] for t in msg[1] if t in states}
            for task in msg[1]:
                if task not in states:
                    states[task], r = _cpu_setup(task)
                    rates.append(r)
            conn.send(rates)
        elif msg[0] == "step":
            n_pl, it_max = 0, 0
            for (start, n), st in states.items():
                out = ss.integrate(synth.strain(n, SEED, AMP, KINC, KINC, start=start), st, PROPS)
                n_pl += int(out["flag"].sum())
                it_max = max(it_max, int(out["n_iter"].max()))
            conn.send((n_pl, it_max))


class CpuArm:
    """numpy oracle on `cores` worker processes; each worker owns fixed chunks of the sample so the
    state built during (untimed) preparation is reused by every timed step."""

    def __init__(self, cores, chunk=250_000):
        import multiprocessing as mp

        ctx = mp.get_context("fork")
        self.cores = cores
        self.chunk = chunk
        self.workers = []
        self.npoints = 0
        for _ in range(cores):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_cpu_worker, args=(b,), daemon=True)
            p.start()
            self.workers.append((p, a))

    def prepare(self, chunks_per_core):
        """(re)assign `chunks_per_core` chunks to every worker; returns the median GP/s per core seen
        while building the state"""
        for w, (_, conn) in enumerate(self.workers):
            tasks = [((w * chunks_per_core + c) * self.chunk, self.chunk) for c in range(chunks_per_core)]
            conn.send(("prepare", tasks))
        rates = []
        for _, conn in self.workers:
            rates += conn.recv()
        self.npoints = self.cores * chunks_per_core * self.chunk
        return statistics.median(rates) if rates else 0.0

    def step(self):
        """one pass of the hot path over the sample; returns (seconds wall, points, per-worker stats)"""
        t0 = time.perf_counter()
        for _, conn in self.workers:
            conn.send(("step",))
        res = [conn.recv() for _, conn in self.workers]
        return time.perf_counter() - t0, self.npoints, res

    def close(self):
        for p, conn in self.workers:
            conn.send(("stop",))
        for p, _ in self.workers:
            p.join(timeout=10)


def c_port_rate(cores, points_per_core=400_000, passes=3):
    """The plain-C restatement of the same update (oracle/c, gcc -O2 -ffp-contract=off, bit-identical to the numpy
    oracle) on `cores` threads: one pass at increment KINC from the state after KINC - 1 increments.  Reported next
    to the numpy figure because a compiled CPU path (the reference's JAX-CPU back-end is one) sits between the two."""
    from oracle import cport
    from oracle import small_strain as ss
    from oracle import synth

    n = points_per_core * cores
    cport.set_threads(cores)
    st = ss.zero_state(n)
    for k in range(1, KINC):
        st = ss.advance(cport.small_strain(synth.strain(n, SEED, AMP, k, KINC), st, PROPS))
    eps = synth.strain(n, SEED, AMP, KINC, KINC)
    best = None
    for _ in range(passes):
        t0 = time.perf_counter()
        out = cport.small_strain(eps, st, PROPS)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": n / best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {n} points of the same workload, plain-C oracle on {cores} threads, best of {passes} passes at increment {KINC}/{KINC}",
            "sample_plastic_fraction": float(out["flag"].mean())}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the CPU arm runs once per box
    cores = os.cpu_count() or 1
    arm = CpuArm(cores)
    # size the sample from a short untimed probe so that steps+warmup stay within ~2 minutes
    rate = arm.prepare(1)  # GP/s per core, measured while building the state
    budget = min(3.0, 100.0 / max(1, args.steps + args.warmup))
    cpc = max(1, min(8, int(rate * budget / arm.chunk)))
    if cpc > 1:
        arm.prepare(cpc)
    for _ in range(args.warmup):
        arm.step()
    t = 0.0
    pts = 0
    for _ in range(args.steps):
        w, p, _ = arm.step()
        t += w
        pts += p
    arm.close()
    try:
        c_port = c_port_rate(cores)
    except Exception as e:  # noqa: BLE001 - the C figure is an extra, the numpy arm is the line's value
        c_port = {"unavailable": str(e)[:200]}
    value = pts / t
    sample = f"{pts // args.steps} points/step ({cores} procs x {cpc} chunks x {arm.chunk}), numpy oracle, increment {KINC}/{KINC} from the state after {KINC - 1} increments"
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "c_port": c_port},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference CPU path = numpy port of the reference algorithm (oracle/); the reference's own jaxmat/JAX back-end is not installable offline (DESIGN.md)",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampling (NVML) during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        self.samples, self.reasons, self.power = [], set(), []
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001 - clocks are best effort, the bench number is not
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {
            "sm_mhz": statistics.median(self.samples),
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "power_w_max": max(self.power) if self.power else None,
            "samples": len(self.samples),
        }


# ------------------------------------------------------------------------------------------------
def workload_config(args, world):
    return {
        "workload": "cfg2: 3D small-strain J2 plasticity + Voce hardening, fp64, synthetic proportional strain histories",
        "points_per_gpu": int(args.n),
        "global_points": int(args.n) * world,
        "properties": PROPS,
        "history": f"counter-based recipe seed {SEED}, amp {AMP}, increment {KINC}/{KINC} after {KINC - 1} state updates",
        "l2": "inputs >> L2 (47.2 GB touched per step per GPU), no flush needed",
        "parallelism": f"points sharded over {world} GPU(s), no data-path collective; NCCL all-reduce of 4 statistics per step",
        "e2e_points_per_gpu": int(args.e2e_n),
    }


def bind_to_gpu_numa_node(index):
    """Pin this rank's host threads (and therefore its first-touch page-locked buffers) to the CPUs NVML
    reports as local to its GPU: with 8 ranks on one box the e2e leg is bound by host-side D2H bandwidth,
    and remote-NUMA pinned buffers halve it.  Best effort."""
    try:
        import pynvml

        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
        return True
    except Exception:  # noqa: BLE001
        return False


def load_traffic(n):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
    (profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum at a stated n), scaled to
    this run's n when the capture was taken at another size; None when no capture is recorded."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)["dxm_small_strain_kernel"]
        per_gp = (t["dram_bytes_read"] + t["dram_bytes_write"]) / t["n"]
        return per_gp * n, f"ncu capture at n={t['n']:.0f} ({t['source']}): {per_gp:.1f} B/point"
    except Exception:  # noqa: BLE001
        return None, None


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


def run_ours(args):
    import torch
    import torch.distributed as dist

    import dolfinx_materials_b200 as jm
    from dolfinx_materials_b200 import _lib, build
    from dolfinx_materials_b200.distributed import allreduce_stats, shard_start
    from dolfinx_materials_b200.material import PinnedArray

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # CPU-baseline workers are forked before this process touches CUDA
    arm = CpuArm(os.cpu_count() or 1) if (rank == 0 and world == 1 and not args.no_cpu) else None
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this framework has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    bind_to_gpu_numa_node(local)
    # stdout carries exactly ONE JSON line: everything else that writes to file descriptor 1 -- NCCL prints its
    # "NCCL version ..." banner there from C, whatever NCCL_DEBUG_FILE says -- is sent to stderr for the whole run;
    # the result line goes to a private duplicate of the original stdout (emit()).
    sys.stdout.flush()
    out_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(out_fd, (json.dumps(obj) + "\n").encode())

    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    build.build_library()
    lib = _lib.load()

    n = int(args.n)
    beh = jm.vonMisesIsotropicHardening(
        elasticity=jm.LinearElasticIsotropic(E=PROPS["E"], nu=PROPS["nu"]),
        yield_stress=jm.VoceHardening(sig0=PROPS["sig0"], sigu=PROPS["sigu"], b=PROPS["b"]),
    )
    m = jm.CUDAMaterial(beh, device=local)
    m.set_data_manager(n)
    stream = torch.cuda.current_stream()
    m.set_stream(stream.cuda_stream)
    start = shard_start(n, rank)

    # ---- load history up to the last increment (untimed) ----------------------------------------
    for k in range(1, KINC):
        m.synth_gradients(SEED, AMP, k, KINC, start=start)
        m.integrate_resident()
        m.data_manager.update()
    m.synth_gradients(SEED, AMP, KINC, KINC, start=start)

    def step():
        s = m.integrate_resident()
        return allreduce_stats(s) if world > 1 else s

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    launches0 = lib.dxm_launch_count()
    kernel_ms = []
    ev0.record(stream)
    for _ in range(args.steps):
        s = step()
        kernel_ms.append(m.last_stats.kernel_ms)
    ev1.record(stream)
    barrier()
    launches = lib.dxm_launch_count() - launches0
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_total, sum(kernel_ms) / len(kernel_ms)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, kms = t.tolist()
    else:
        kms = sum(kernel_ms) / len(kernel_ms)
    ms_per_step = ms_total / args.steps
    value = world * n / (ms_per_step * 1e-3)

    # ---- end to end through the reference-facing call (host buffers) ---------------------------
    ne = int(min(args.e2e_n, n))
    e2e = None
    if ne > 0:
        me = jm.CUDAMaterial(beh, device=local)
        me.set_data_manager(ne)
        # same history on the e2e points: state after 3 increments, then the timed call repeats increment 4
        for k in range(1, KINC):
            me.synth_gradients(SEED, AMP, k, KINC, start=start)
            me.integrate_resident()
            me.data_manager.update()
        me.synth_gradients(SEED, AMP, KINC, KINC, start=start)
        me.integrate_resident()
        grads = PinnedArray((ne, 6))
        grads.array[...] = me.device_view("strain").T.cpu().numpy()
        e_steps = max(2, min(args.steps, 5))
        me.integrate(grads.array)  # warm-up: allocates staging + pinned outputs
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            flux, isv, ct = me.integrate(grads.array)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = t.item()
        e2e = {
            "value": world * ne * e_steps / dt,
            "unit": UNIT,
            "h2d_bytes_per_step": ne * 6 * 8,
            "d2h_bytes_per_step": ne * (6 + 7 + 36) * 8,
            "points_per_gpu": ne,
            "steps": e_steps,
            "ms_per_step": 1e3 * dt / e_steps,
            "api": "CUDAMaterial.integrate(host (n,6) gradients) -> host (flux, isv, Ct), pinned buffers",
        }
        # the e2e outputs must equal the resident results bit for bit
        ok = bool(np.array_equal(flux, me.device_view("stress").T.cpu().numpy()))
        e2e["matches_resident"] = ok
        del me

    # ---- CPU baseline + parity spot check on rank 0 at N = 1 -------------------------------------
    cpu = None
    if arm is not None:
        cores = arm.cores
        rate = arm.prepare(1)
        cpc = max(1, min(4, int(rate * 1.5 / arm.chunk)))
        if cpc > 1:
            arm.prepare(cpc)
        arm.step()
        wall, pts, res = arm.step()
        arm.close()
        cpu = {
            "value": pts / wall,
            "unit": UNIT,
            "cores": cores,
            "kind": "port",
            "sample": f"first {pts} points of the same workload ({cores} procs x {cpc} chunks x {arm.chunk}), numpy oracle, one pass at increment {KINC}/{KINC}",
            "sample_plastic_fraction": sum(r[0] for r in res) / pts,
        }
        try:
            cpu["c_port"] = c_port_rate(cores)
        except Exception as e:  # noqa: BLE001
            cpu["c_port"] = {"unavailable": str(e)[:200]}

    if rank == 0:
        peak, peak_src = load_peaks()
        achieved = BYTES_PER_GP * n / (kms * 1e-3) / 1e9
        traffic, traffic_src = load_traffic(n)
        line = {
            "metric": METRIC,
            "value": value,
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(args, world),
            "roofline": {
                "bound": "hbm",
                "achieved": achieved,
                "peak": peak,
                "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": traffic,
                "traffic_source": traffic_src,
                "kernel": "dxm_small_strain_kernel<HARD_GENERAL,uniform,PPT=1>",
                "kernel_ms": kms,
                "algorithmic_bytes_per_launch": BYTES_PER_GP * n,
                "moved_bytes_per_launch": BYTES_MOVED_PER_GP * n,
                "moved_gbs": BYTES_MOVED_PER_GP * n / (kms * 1e-3) / 1e9,
                "moved_frac": BYTES_MOVED_PER_GP * n / (kms * 1e-3) / 1e9 / peak,
                "peak_source": peak_src,
            },
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "stats": {"plastic_fraction": s.n_plastic / (n * world), "n_fail": s.n_fail, "max_iter": s.max_iter,
                      "max_residual": s.max_residual},
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=float, default=1e8, help="Gauss points per GPU")
    ap.add_argument("--e2e-n", type=float, default=1e7, help="Gauss points per GPU for the host-buffer e2e leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
