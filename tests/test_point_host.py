"""The J2 and FeFp kernels' per-point routines (``csrc/dxm_small_strain.cuh``: ``j2_point`` / ``j2_tangent_entry`` /
``point_props``; ``csrc/dxm_fefp.cuh``: ``fefp_point``, all ``__host__ __device__``) executed on the CPU and compared
bit for bit with the oracle -- the same code the GPU runs per Gauss point, checked where no GPU is available.
(The GPU parity tests proper are ``tests/test_small_strain_gpu.py``, ``test_fefp_gpu.py``, ``test_table_hardening_gpu.py``.)"""

import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import fefp
from oracle import small_strain as ss
from oracle import synth

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "point_host_check.cu")
LIB = os.path.join(HERE, "_build", "libpoint_host_check.so")
HARD_NONE, HARD_LINEAR, HARD_GENERAL, HARD_TABLE = 0, 1, 2, 3  # dxm::Hardening


@pytest.fixture(scope="module")
def host():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    deps = [SRC] + [os.path.join(HERE, "..", "dolfinx_materials_b200", "csrc", f)
                    for f in ("dxm_fefp.cuh", "dxm_small_strain.cuh", "dxm_canon.cuh")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-fmad=false",
                        "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-o", LIB, SRC], check=True)
    return ctypes.CDLL(LIB)


def c(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def prop_rows(props, n):
    """(E, nu, sig0, H, sigu, b) as the library stores them: 6 scalars, or [6][n] rows if any entry is per point."""
    vals = [props["E"], props["nu"], props.get("sig0", np.inf), props.get("H", 0.0),
            props.get("sigu", props.get("sig0", np.inf)), props.get("b", 0.0)]
    perpoint = any(np.ndim(v) > 0 for v in vals)
    if perpoint:
        return np.ascontiguousarray(np.stack([np.broadcast_to(np.asarray(v, dtype=float), (n,)) for v in vals])), 1
    return np.array(vals, dtype=float), 0


def run_j2(lib, eps, st, props, hard, vote=1):
    n = eps.shape[0]
    eps = np.ascontiguousarray(eps)
    e_old, s_old = np.ascontiguousarray(st["strain"]), np.ascontiguousarray(st["stress"])
    p_old, ep_old = np.ascontiguousarray(st["p"]).reshape(n), np.ascontiguousarray(st["epsp"])
    rows, perpoint = prop_rows(props, n)
    if "table" in props:
        table = np.ascontiguousarray(np.stack(ss.table_slopes(*props["table"])))
    else:
        table = np.zeros((3, 1))
    sig, p, epsp, ct = np.empty((n, 6)), np.empty(n), np.empty((n, 6)), np.empty((n, 6, 6))
    flag, fail, it, rs = np.empty(n, np.uint8), np.empty(n, np.uint8), np.empty(n, np.int32), np.empty(n)
    lib.j2_host(ctypes.c_int64(n), c(eps), c(e_old), c(s_old), c(p_old), c(ep_old), c(rows), ctypes.c_int(perpoint),
                ctypes.c_int(hard), c(table), ctypes.c_int(table.shape[1]), ctypes.c_int(vote), c(sig), c(p), c(epsp), c(ct),
                c(flag), c(it), c(rs), c(fail))
    return {"strain": eps, "stress": sig, "p": p, "epsp": epsp, "Ct": ct, "flag": flag, "n_iter": it, "resid": rs, "fail": fail}


def assert_same(got, ref, keys, ctx=()):
    """Bitwise equality; NaNs must sit at the same places (their sign / payload bits are not specified by IEEE 754 and
    differ between x86 SSE, numpy's constants and the GPU's canonical NaN)."""
    for key in keys:
        g, r = got[key], np.ascontiguousarray(ref[key])
        assert g.shape == r.shape and g.dtype == r.dtype, (key, *ctx)
        if g.dtype.kind == "f":
            nan = np.isnan(r)
            assert np.array_equal(np.isnan(g), nan), (key, *ctx)
            g, r = np.where(nan, 0.0, g), np.where(nan, 0.0, r)
        assert g.tobytes() == r.tobytes(), (key, *ctx)


J2_KEYS = ("flag", "n_iter", "fail", "stress", "p", "epsp", "Ct", "resid")
CASES = {
    "elastic": (dict(E=70e3, nu=0.3, sig0=np.inf), HARD_NONE),
    "linear": (dict(E=70e3, nu=0.3, sig0=250.0, H=700.0), HARD_LINEAR),
    "perfect": (dict(E=70e3, nu=0.3, sig0=250.0, H=0.0), HARD_LINEAR),
    "voce": (dict(E=70e3, nu=0.3, sig0=350.0, sigu=500.0, b=1e3), HARD_GENERAL),
    "voce+linear": (dict(E=210e3, nu=0.25, sig0=400.0, sigu=650.0, b=40.0, H=1500.0), HARD_GENERAL),
    "general-without-saturation": (dict(E=70e3, nu=0.3, sig0=250.0, H=700.0), HARD_GENERAL),  # closed form inside GENERAL
    "table": (dict(E=70e3, nu=0.3, table=(np.array([0.0, 2e-4, 1e-3, 4e-3, 2e-2, 0.2]),
                                           np.array([300.0, 340.0, 390.0, 430.0, 470.0, 520.0]))), HARD_TABLE),
}


@pytest.mark.parametrize("case", list(CASES))
def test_j2_point_routine_equals_oracle_bit_for_bit(host, case):
    props, hard = CASES[case]
    n, K = 20000, 4
    st = ss.zero_state(n)
    for k in range(1, K + 1):
        eps = synth.strain(n, 3, 1.25e-2, k, K)
        ref = ss.integrate(eps, st, props)
        for vote in (1, 0):
            assert_same(run_j2(host, eps, st, props, hard, vote), ref, J2_KEYS, (case, k, vote))
        st = ss.advance(ref)
    assert ref["fail"].sum() == 0
    if hard != HARD_NONE:
        assert 0.3 < ref["flag"].mean() < 0.95
    if case in ("voce", "voce+linear"):
        assert ref["n_iter"].max() >= 3
    if case == "table":
        assert ref["n_iter"].max() >= 2  # segment crossings


def test_j2_point_routine_per_point_properties(host):
    """Heterogeneous batch (config 4: a different law per point behind one handle): the general instantiation with
    per-point rows, including elastic points (sig0 = inf) and linear points (sigu = sig0)."""
    rng = np.random.default_rng(7)
    n = 12000
    kind = rng.integers(0, 3, n)
    sig0 = np.where(kind == 0, np.inf, rng.uniform(200.0, 400.0, n))
    sigu = np.where(kind == 2, sig0 + rng.uniform(50.0, 200.0, n), sig0)
    props = dict(E=rng.uniform(60e3, 210e3, n), nu=rng.uniform(0.1, 0.4, n), sig0=sig0, sigu=sigu,
                 H=np.where(kind == 1, rng.uniform(0.0, 2e3, n), 0.0), b=np.where(kind == 2, rng.uniform(10.0, 2e3, n), 0.0))
    st = ss.zero_state(n)
    for k in range(1, 4):
        eps = synth.strain(n, 11, 1e-2, k, 3)
        ref = ss.integrate(eps, st, props)
        assert_same(run_j2(host, eps, st, props, HARD_GENERAL), ref, J2_KEYS, (k,))
        st = ss.advance(ref)
    assert ref["fail"].sum() == 0 and ref["flag"][kind == 0].sum() == 0 and ref["flag"][kind > 0].mean() > 0.3


def test_j2_point_routine_extreme_inputs(host):
    """NaN / inf strains raise the fail flag exactly as in the oracle; zero increment, pure volumetric and huge steps."""
    props, hard = CASES["voce"]
    rows = [[0, 0, 0, 0, 0, 0], [1e-2, 1e-2, 1e-2, 0, 0, 0], [np.nan, 0, 0, 0, 0, 0], [np.inf, 0, 0, 0, 0, 0],
            [0.5, -0.2, 0.1, 0.3, -0.4, 0.2], [1e-3, -5e-4, -5e-4, 0, 0, 0], [5.0, -2.5, -2.5, 0, 0, 0]]
    eps = np.array(rows, dtype=float)
    st = ss.zero_state(len(rows))
    ref = ss.integrate(eps, st, props)
    assert_same(run_j2(host, eps, st, props, hard), ref, J2_KEYS)
    assert ref["fail"][2] == 1 and ref["fail"][3] == 1 and ref["fail"][[0, 1, 4, 5, 6]].sum() == 0


def test_j2_point_routine_newton_cap(host):
    """A local solve that cannot converge (non-monotone law: negative saturation with a huge rate) stops at the cap
    with the fail flag -- same iteration count and residual as the oracle."""
    props = dict(E=70e3, nu=0.3, sig0=300.0, sigu=-1e7, b=1e5)
    n = 256
    eps = synth.strain(n, 5, 2e-2, 1, 1)
    st = ss.zero_state(n)
    ref = ss.integrate(eps, st, props)
    assert_same(run_j2(host, eps, st, props, HARD_GENERAL), ref, J2_KEYS)
    assert ref["fail"].sum() > 0 and ref["n_iter"].max() == 25


# ---- FeFp ----------------------------------------------------------------------------------------------------------
FEFP_PROPS = dict(E=70e3, nu=0.3, sig0=500.0, sigu=750.0, b=1000.0)
FEFP_KEYS = ("flag", "n_iter", "fail", "PK1", "p", "be_bar", "Ct", "resid")


def run_fefp(lib, F, st, props, vote=1):
    n = F.shape[0]
    soa = lambda a, d: np.ascontiguousarray(np.asarray(a, dtype=float).reshape(n, d).T)  # noqa: E731
    Fs, Fo, po, beo = soa(F, 9), soa(st["F"], 9), np.ascontiguousarray(st["p"]).reshape(n), soa(st["be_bar"], 6)
    rows, perpoint = prop_rows(props, n)
    P, p, be, ct = np.empty((9, n)), np.empty(n), np.empty((6, n)), np.empty((81, n))
    flag, fail, it, rs = np.zeros(n, np.uint8), np.zeros(n, np.uint8), np.zeros(n, np.int32), np.zeros(n)
    npl, nf = ctypes.c_uint64(0), ctypes.c_uint64(0)
    lib.fefp_host(ctypes.c_int64(n), c(Fs), c(Fo), c(po), c(beo), c(rows), ctypes.c_int(perpoint), ctypes.c_int(vote), c(P), c(p),
                  c(be), c(ct), c(flag), c(it), c(rs), c(fail), ctypes.byref(npl), ctypes.byref(nf))
    assert npl.value == int(flag.sum()) and nf.value == int(fail.sum())  # the statistics the kernel reduces
    return {"PK1": np.ascontiguousarray(P.T), "p": p, "be_bar": np.ascontiguousarray(be.T),
            "Ct": np.ascontiguousarray(ct.T).reshape(n, 9, 9), "flag": flag, "n_iter": it, "resid": rs, "fail": fail}


def test_fefp_point_routine_reference_test_script(host, Nbatch=10):
    """The script of the reference's tests/test_FeFp_jax.py:6-33 through the kernel's point routine on the CPU."""
    eps, Nsteps = 2e-2, 20
    st = fefp.virgin_state(Nbatch)
    for t in np.linspace(0, 1.0, Nsteps)[1:]:
        F = np.zeros((Nbatch, 9))
        F[:, 0] = 1 + eps * t
        F[:, [1, 2]] = 1 - eps / 2 * t
        ref = fefp.integrate(F, st, FEFP_PROPS)
        assert_same(run_fefp(host, F, st, FEFP_PROPS), ref, FEFP_KEYS, (t,))
        st = fefp.advance(ref)
    assert abs(ref["p"][0] - 1.076097e-2) < 5e-9 and abs(ref["PK1"][0, 0] - 473.1527) < 5e-5


@pytest.mark.parametrize("seed", [0, 1])
def test_fefp_point_routine_random_history(host, seed):
    n, K = 8000, 4
    st = fefp.virgin_state(n)
    for k in range(1, K + 1):
        F = synth.defgrad(n, seed, 4e-2, k, K)
        ref = fefp.integrate(F, st, FEFP_PROPS)
        for vote in (1, 0):
            assert_same(run_fefp(host, F, st, FEFP_PROPS, vote), ref, FEFP_KEYS, (k, vote))
        st = fefp.advance(ref)
    assert ref["fail"].sum() == 0 and 0.2 < ref["flag"].mean() < 0.98 and ref["n_iter"].max() >= 3


def test_fefp_point_routine_per_point_properties_and_bad_input(host):
    rng = np.random.default_rng(3)
    n = 4000
    props = dict(E=rng.uniform(60e3, 210e3, n), nu=rng.uniform(0.1, 0.4, n), sig0=rng.uniform(200.0, 600.0, n),
                 H=rng.uniform(0.0, 1e3, n), b=rng.uniform(10.0, 1e3, n))
    props["sigu"] = props["sig0"] + rng.uniform(0.0, 300.0, n)
    st = fefp.virgin_state(n)
    for k in range(1, 3):
        F = synth.defgrad(n, 9, 3e-2, k, 2)
        if k == 2:
            F[5, 0] = np.nan  # non-finite gradient -> fail flag, NaN outputs, identical bits
            F[6, :] = 0.0     # singular F
        ref = fefp.integrate(F, st, props)
        assert_same(run_fefp(host, F, st, props), ref, FEFP_KEYS, (k,))
        st = fefp.advance(ref)
    assert ref["fail"][5] == 1 and ref["fail"][6] == 1 and ref["fail"].sum() == 2


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_point_routines_differential_fuzz_over_extreme_regimes(host, seed):
    """Strain amplitudes from 1e-14 to 30, moduli over 8 decades, Poisson ratios from -0.9 to 0.499, hardening rates
    over 10 decades: whatever regime a point ends up in (elastic, converged, capped, non-finite) the kernel routines
    and the oracle agree bit for bit -- flags, iteration counts, residuals, results."""
    rng = np.random.default_rng(seed)
    idx9 = [(0, 0), (1, 1), (2, 2), (0, 1), (1, 0), (0, 2), (2, 0), (1, 2), (2, 1)]
    regimes = set()
    for _ in range(4):
        n = 3000
        props = dict(E=10 ** rng.uniform(0, 8), nu=rng.uniform(-0.9, 0.499), sig0=10 ** rng.uniform(-3, 5),
                     b=10 ** rng.uniform(-3, 7), H=float(rng.choice([0.0, 10 ** rng.uniform(-6, 6)])))
        props["sigu"] = props["sig0"] * rng.uniform(0.5, 3.0)
        eps = rng.standard_normal((n, 6)) * 10.0 ** rng.uniform(-14, 1.5, (n, 1))
        st = ss.zero_state(n)
        with np.errstate(all="ignore"):
            for k in (1, 2):
                ref = ss.integrate(eps * k / 2, st, props)
                assert_same(run_j2(host, eps * k / 2, st, props, HARD_GENERAL), ref, J2_KEYS, ("j2", props, k))
                st = ss.advance(ref)
            regimes |= {("j2", int(f), int(x)) for f, x in zip(ref["flag"], ref["fail"])}
            G = rng.standard_normal((n, 3, 3)) * 10.0 ** rng.uniform(-12, -0.3, (n, 1, 1))
            F = np.ascontiguousarray(np.stack([(np.eye(3) + G)[:, i, j] for (i, j) in idx9], axis=1))
            st = fefp.virgin_state(n)
            for Fk in (np.ascontiguousarray(0.5 * (st["F"] + F)), F):
                ref = fefp.integrate(Fk, st, props)
                assert_same(run_fefp(host, Fk, st, props), ref, FEFP_KEYS, ("fefp", props))
                st = fefp.advance(ref)
            regimes |= {("fefp", int(f), int(x)) for f, x in zip(ref["flag"], ref["fail"])}
    assert {("j2", 0, 0), ("j2", 1, 0), ("fefp", 0, 0), ("fefp", 1, 0)} <= regimes
