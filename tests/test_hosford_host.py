"""The Hosford kernel's per-point routine (``csrc/dxm_hosford.cuh``: ``hosford_point``, ``__host__ __device__``)
executed on the CPU and compared bit for bit with the oracle -- the same code path the GPU runs per Gauss point,
checked where no GPU is available.  (The GPU parity tests proper are in ``tests/test_hosford_gpu.py``.)"""

import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import hosford as ho
from oracle import small_strain as ss
from oracle import synth

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hosford_host_check.cu")
LIB = os.path.join(HERE, "_build", "libhosford_host_check.so")


@pytest.fixture(scope="module")
def host():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    deps = [SRC] + [os.path.join(HERE, "..", "dolfinx_materials_b200", "csrc", f)
                    for f in ("dxm_hosford.cuh", "dxm_small_strain.cuh", "dxm_canon.cuh")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-fmad=false",
                        "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-o", LIB, SRC], check=True)
    return ctypes.CDLL(LIB)


def run(lib, eps, st, props, split=0, generic=0):
    n = eps.shape[0]
    c = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    eps = np.ascontiguousarray(eps)
    e_old, s_old = np.ascontiguousarray(st["strain"]), np.ascontiguousarray(st["stress"])
    p_old, ep_old = np.ascontiguousarray(st["p"]).reshape(n), np.ascontiguousarray(st["epsp"])
    sig, p, epsp, ct = np.empty((n, 6)), np.empty(n), np.empty((n, 6)), np.empty((n, 6, 6))
    flag, fail, it, rs = np.empty(n, np.uint8), np.empty(n, np.uint8), np.empty(n, np.int32), np.empty(n)
    ncand = ctypes.c_int64(0)
    a = props["a"]
    bound = (2.0 ** (a - 1) + 1.0) ** (1.0 / a) / np.sqrt(3.0) * (1.0 + 1e-9)  # hosford_bound(), dxm_hosford_api.cu
    lib.hosford_host(ctypes.c_int64(n), c(eps), c(e_old), c(s_old), c(p_old), c(ep_old), ctypes.c_double(props["E"]),
                     ctypes.c_double(props["nu"]), ctypes.c_double(props["sig0"]), ctypes.c_double(props.get("H", 0.0)),
                     ctypes.c_double(props.get("sigu", props["sig0"])), ctypes.c_double(props.get("b", 0.0)),
                     ctypes.c_int(props["a"]), ctypes.c_double(bound), c(sig), c(p), c(epsp), c(ct), c(flag), c(it), c(rs),
                     c(fail), ctypes.c_int(split), ctypes.byref(ncand), ctypes.c_int(generic))
    return {"candidates": ncand.value, "strain": eps, "stress": sig, "p": p, "epsp": epsp, "Ct": ct, "flag": flag, "n_iter": it, "resid": rs,
            "fail": fail}


@pytest.mark.parametrize("a", [2, 4, 6, 8, 10, 20])
def test_kernel_point_routine_equals_oracle_bit_for_bit(host, a):
    props = dict(E=70e3, nu=0.3, sig0=200.0, H=10.0, a=a)
    n = 20000
    st = ss.zero_state(n)
    for k in range(1, 4):
        eps = synth.strain(n, a, 1.25e-2, k, 3)
        ref = ho.integrate(eps, st, props)
        for split, generic in ((0, 0), (1, 0), (1, 1)):  # unrolled-exponent and generic instantiations
            got = run(host, eps, st, props, split, generic)
            for key in ("flag", "n_iter", "fail", "stress", "p", "epsp", "Ct", "resid"):
                assert np.array_equal(got[key], ref[key]), (key, k, split, generic)
        # the light pass hands over every plastic point and only a thin shell of elastic ones near the surface
        assert ref["flag"].sum() <= got["candidates"] <= ref["flag"].sum() + 0.2 * n
        st = ss.advance(ref)
    assert 0.3 < ref["flag"].mean() < 0.95 and ref["fail"].sum() == 0


def test_kernel_point_routine_degenerate_and_extreme_inputs(host):
    props = dict(E=70e3, nu=0.3, sig0=200.0, H=1e-6, a=10)
    rows = [[0, 0, 0, 0, 0, 0], [8e-3, -3e-3, -3e-3, 0, 0, 0], [1e-2, 1e-2, 1e-2, 0, 0, 0], [0, 0, 0, 2e-2, 0, 0],
            [0.3, -0.1, 0.05, 0.2, -0.3, 0.1], [np.nan, 0, 0, 0, 0, 0], [1e-3, 1e-3, -2e-3, 0, 0, 0]]
    eps = np.array(rows, dtype=float)
    st = ss.zero_state(len(rows))
    ref = ho.integrate(eps, st, props)
    for split in (0, 1):
        got = run(host, eps, st, props, split)
        for key in ("flag", "n_iter", "fail"):
            assert np.array_equal(got[key], ref[key]), key
        for key in ("stress", "p", "epsp", "Ct", "resid"):
            assert got[key].tobytes() == ref[key].tobytes(), key  # bitwise, NaN included
    assert ref["fail"][5] == 1 and ref["fail"].sum() == 1


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_kernel_point_routine_random_states_and_properties(host, seed):
    """Arbitrary admissible previous states (pre-stress, accumulated plastic strain) and properties, non-proportional
    increments, unloading: kernel routine == oracle bit for bit, fused and split."""
    rng = np.random.default_rng(seed)
    n = 4000
    for a in (6, 8, 10, 12):
        props = dict(E=float(rng.uniform(50e3, 210e3)), nu=float(rng.uniform(0.05, 0.45)),
                     sig0=float(rng.uniform(100.0, 400.0)), H=float(rng.choice([0.0, 10.0, 2e3])), a=a)
        st = ss.zero_state(n)
        # build a non-trivial state with two oracle steps in unrelated directions, then test a third one (some
        # points reload, some unload, some stay elastic)
        for k in range(2):
            st = ss.advance(ho.integrate(synth.strain(n, 10 * seed + k, 8e-3, 1, 1), st, props))
        eps = st["strain"] + rng.standard_normal((n, 6)) * rng.uniform(0, 2e-3, (n, 1))
        ref = ho.integrate(eps, st, props)
        assert ref["fail"].sum() == 0 and 0.02 < ref["flag"].mean() < 0.98
        for split in (0, 1):
            got = run(host, eps, st, props, split)
            for key in ("flag", "n_iter", "fail", "stress", "p", "epsp", "Ct", "resid"):
                assert np.array_equal(got[key], ref[key]), (key, a, split)


@pytest.mark.parametrize("a", [2, 6, 10, 14])
def test_kernel_point_routine_voce_hardening(host, a):
    """General hardening law (Voce term + linear term) behind the Hosford criterion: kernel routine == oracle."""
    props = dict(E=70e3, nu=0.3, sig0=350.0, sigu=500.0, b=1e3, H=25.0, a=a)
    n = 20000
    st = ss.zero_state(n)
    for k in range(1, 4):
        eps = synth.strain(n, a, 1.25e-2, k, 3)
        ref = ho.integrate(eps, st, props)
        for split in (0, 1):
            got = run(host, eps, st, props, split)
            for key in ("flag", "n_iter", "fail", "stress", "p", "epsp", "Ct", "resid"):
                assert np.array_equal(got[key], ref[key]), (key, k, split)
        st = ss.advance(ref)
    assert 0.3 < ref["flag"].mean() < 0.95 and ref["fail"].sum() == 0


def test_kernel_point_routine_fuzz_over_regimes(host):
    """Properties over orders of magnitude, strain steps from 1e-3 to 300 x the yield strain, repeated eigenvalues,
    nearly hydrostatic and nearly uniaxial states, exponents 2..64, with and without the Voce term, load reversal:
    no failed local solve, at most a dozen Newton iterations, kernel routine == oracle bit for bit (NaN-safe compare)."""
    rng = np.random.default_rng(123)
    n = 1500
    worst = 0
    for trial in range(24):
        a = int(rng.choice([2, 4, 6, 8, 10, 12, 16, 20, 32, 64]))
        props = dict(E=float(10 ** rng.uniform(3, 6)), nu=float(rng.uniform(0.0, 0.495)), sig0=float(10 ** rng.uniform(0, 3)),
                     H=float(rng.choice([0.0, 1e-6, 10.0, 1e4, 1e6])), a=a)
        if rng.random() < 0.4:
            props.update(sigu=props["sig0"] * float(rng.uniform(1.0, 3.0)), b=float(10 ** rng.uniform(0, 4)))
        scale = props["sig0"] / props["E"] * 10 ** rng.uniform(-3, 2.5, size=(n, 1))
        d = rng.standard_normal((n, 6))
        if trial % 4 == 1:  # repeated eigenvalues
            d[:, 3:] = 0
            d[:, 1] = d[:, 2]
        if trial % 4 == 2:  # nearly hydrostatic
            d[:, :3] = d[:, :1] + 1e-9 * d[:, :3]
            d[:, 3:] *= 1e-9
        if trial % 4 == 3:  # nearly uniaxial strain
            d[:, 1:] *= 1e-7
        st = ss.zero_state(n)
        for k in range(2):
            eps = st["strain"] + scale * d * (1 if k == 0 else rng.uniform(-1, 1, size=(n, 1)))
            ref = ho.integrate(eps, st, props)
            got = run(host, eps, st, props, int(rng.integers(0, 2)))
            for key in ("flag", "n_iter", "fail", "stress", "p", "epsp", "Ct", "resid"):
                assert got[key].tobytes() == ref[key].tobytes(), (key, trial, props)
            assert ref["fail"].sum() == 0 and np.isfinite(ref["Ct"]).all(), (trial, props)
            worst = max(worst, int(ref["n_iter"].max()))
            st = ss.advance(ref)
    assert worst <= 14
