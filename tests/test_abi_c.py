"""include/dxm.h from plain C: the header compiles as C99 with -Wall -Wextra -Werror, a C program links against
libdxm_cuda.so and exercises the entry points that need no GPU (CPU test) and a small update through dxm_integrate
(GPU test) -- the C-ABI boundary used without Python."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "abi_c", "abi_smoke.c")
EXE = os.path.join(ROOT, "tests", "_build", "abi_smoke")


@pytest.fixture(scope="module")
def exe(jm):
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("gcc not available")
    libdir = os.path.join(ROOT, "dolfinx_materials_b200", "lib")
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE,
                    "-L", libdir, "-ldxm_cuda", "-lm", f"-Wl,-rpath,{libdir}"], check=True)
    return EXE


def test_header_is_c99_and_host_entry_points_work_from_c(exe):
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "ok" in r.stdout


@pytest.mark.gpu
def test_update_through_the_c_abi_from_c(exe):
    r = subprocess.run([exe, "gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    assert "closed form matched" in r.stdout
