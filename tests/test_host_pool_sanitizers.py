"""The library's host thread pool (``csrc/dxm_host_mirror.hpp``: packed-tangent mirror, row gather / scatter -- the host
passes of the cell-subset exchange) under ThreadSanitizer and AddressSanitizer + UBSan, driven from three caller threads
at once (several material handles may be used from different threads of one process)."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_pool_check.cpp")


@pytest.mark.parametrize("sanitizer", ["thread", "address,undefined"])
def test_host_pool_is_clean_under_sanitizers(sanitizer, tmp_path):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path / "pool_check")
    build = subprocess.run([gxx, "-O1", "-g", "-std=c++17", f"-fsanitize={sanitizer}", "-o", exe, SRC, "-lpthread"],
                           capture_output=True, text=True)
    if build.returncode != 0 and "sanitize" in build.stderr:
        pytest.skip(f"-fsanitize={sanitizer} not supported by this toolchain")
    assert build.returncode == 0, build.stderr[-2000:]
    env = dict(os.environ, DXM_HOST_THREADS="6", TSAN_OPTIONS="halt_on_error=1", ASAN_OPTIONS="detect_leaks=0")
    run = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=600)
    assert run.returncode == 0 and "bad=0" in run.stdout, (run.stdout + run.stderr)[-3000:]
    assert "WARNING: ThreadSanitizer" not in run.stderr and "ERROR: AddressSanitizer" not in run.stderr
    assert "runtime error" not in run.stderr
