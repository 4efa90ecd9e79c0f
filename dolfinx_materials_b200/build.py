"""In-tree build of ``lib/libdxm_cuda.so`` with nvcc for sm_100a (the only target)."""

import os
import pathlib
import shutil
import subprocess

ROOT = pathlib.Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
# DXM_UNFUSED=1 selects the A/B build with every hand-placed fused multiply-add split into two roundings again
# (-DDXM_UNFUSED, csrc/dxm_canon.cuh): the round-1 arithmetic -- bit-identical to the committed golden histories and to
# oracle.canon.unfused(), ~40 % more FP64 instructions in the finite-strain update (tests/test_unfused_gpu.py)
UNFUSED = os.environ.get("DXM_UNFUSED", "0") not in ("", "0")
# DXM_VARIANT=<name> [DXM_VARIANT_DEFS="-DA -DB"]: an experiment build lib/libdxm_cuda_<name>.so beside the product
# library (kernel A/B runs on one box: the same variable selects it at load time, _lib.py)
VARIANT = "unfused" if UNFUSED else os.environ.get("DXM_VARIANT", "")
VARIANT_DEFS = os.environ.get("DXM_VARIANT_DEFS", "").split()
LIB = ROOT / "lib" / (f"libdxm_cuda_{VARIANT}.so" if VARIANT else "libdxm_cuda.so")

NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-O3",
    "-std=c++17",
    # the compiler never fuses a product into an addition on its own: the canonical arithmetic places every fused
    # multiply-add by hand (fma_c / fms_c / fnma_c), which is what makes kernel results bit-comparable with the CPU oracle
    "-fmad=false",
    *(["-DDXM_UNFUSED"] if UNFUSED else []),
    *VARIANT_DEFS,
    "-Xcompiler",
    "-fPIC",
]

# translation units of the library (kernels live in the .cuh files they include)
UNITS = ["dxm_api.cu", "dxm_hosford_api.cu", "dxm_fe_api.cu", "dxm_peaks.cu", "dxm_comm.cu"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libdxm_cuda.so")
    return exe


def sources():
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.hpp")) + [ROOT.parent / "include" / "dxm.h"]


def needs_build():
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(s.stat().st_mtime > t for s in sources())


def build_library(force=False, verbose=False):
    """Compile the CUDA library if it is missing or older than its sources."""
    if not force and not needs_build():
        return LIB
    LIB.parent.mkdir(parents=True, exist_ok=True)
    objdir = ROOT / "lib" / (f"obj_{VARIANT}" if VARIANT else "obj")
    objdir.mkdir(exist_ok=True)
    nvcc = _nvcc()
    procs = []
    for unit in UNITS:  # compiled concurrently, one nvcc per translation unit
        cmd = [nvcc, *NVCC_FLAGS, "-c", "-o", str(objdir / (unit[:-3] + ".o")), str(CSRC / unit)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    log = ""
    for cmd, pr in procs:
        out, err = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + out + err)
        log += err
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(LIB),
           *[str(objdir / (u[:-3] + ".o")) for u in UNITS], "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    import sys

    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
