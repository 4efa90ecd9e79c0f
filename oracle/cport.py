"""ctypes front-end of the plain-C oracle (``oracle/c/dxm_oracle.c``), same call signatures and return
dictionaries as ``oracle.small_strain.integrate`` / ``oracle.fefp.integrate``.  TEST INFRASTRUCTURE ONLY."""

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")
LIB_UNFUSED = os.path.join(HERE, "_build", "liboracle_unfused.so")  # every explicit fma split again (round-1 arithmetic)
_lib = None
_lib_unfused = None
_NAMES = ("E", "nu", "sig0", "H", "sigu", "b")


def _cpu_has_fma():
    try:
        with open("/proc/cpuinfo") as f:
            return " fma " in f.read().replace("\n", " ")
    except OSError:
        return True


def load(build=True, fused=None):
    """The C oracle; ``fused=None`` follows ``oracle.canon.FUSED`` (False: the -DDXO_UNFUSED twin)."""
    global _lib, _lib_unfused
    if _lib is None:
        srcs = [os.path.join(HERE, "c", f) for f in os.listdir(os.path.join(HERE, "c"))] + [os.path.join(HERE, "Makefile")]
        stale = any(not os.path.exists(p) or os.path.getmtime(p) < max(os.path.getmtime(s) for s in srcs)
                    for p in (LIB, LIB_UNFUSED))
        if build and stale:
            subprocess.run(["make", "-s", "-C", HERE] + ([] if _cpu_has_fma() else ["FMAFLAG="]), check=True)
        elif not _cpu_has_fma():  # prebuilt with -mfma for a CPU without it: rebuild on libm's fma()
            subprocess.run(["make", "-s", "-B", "-C", HERE, "FMAFLAG="], check=True)
        _lib = ctypes.CDLL(LIB)
        _lib_unfused = ctypes.CDLL(LIB_UNFUSED)
    if fused is None:
        from . import canon

        fused = canon.FUSED
    return _lib if fused else _lib_unfused


THREADS = 1  # row blocks are processed on this many Python threads (ctypes releases the GIL)


def set_threads(n):
    global THREADS
    THREADS = max(1, int(n))


def _run_blocks(n, call):
    """call(lo, hi) over contiguous row blocks, in parallel when THREADS > 1"""
    if THREADS == 1 or n < 4096:
        call(0, n)
        return
    from concurrent.futures import ThreadPoolExecutor

    edges = np.linspace(0, n, THREADS + 1).astype(np.int64)
    with ThreadPoolExecutor(THREADS) as ex:
        list(ex.map(lambda k: call(int(edges[k]), int(edges[k + 1])), range(THREADS)))


def _props(props, n):
    vals = dict(props)
    vals.setdefault("H", 0.0)
    vals.setdefault("sigu", vals["sig0"])
    vals.setdefault("b", 0.0)
    arrs, per = [], []
    for k in _NAMES:
        a = np.ascontiguousarray(np.asarray(vals[k], dtype=np.float64).ravel())
        if a.size not in (1, n):
            raise ValueError(f"property {k}: {a.size} values for {n} points")
        arrs.append(a)
        per.append(1 if a.size == n and n > 1 or (a.size == n and np.asarray(vals[k]).ndim > 0) else 0)
    pers = (ctypes.c_int * 6)(*per)

    def ptrs(lo):
        return (ctypes.c_void_p * 6)(*[a.ctypes.data + (8 * lo if f else 0) for a, f in zip(arrs, per)])

    return arrs, ptrs, pers


def _c(a, lo=0):
    return ctypes.c_void_p(a.ctypes.data + lo * a.strides[0])


def small_strain(eps, state, props, newton_cap=25, rtol=1e-12):
    lib = load()
    eps = np.ascontiguousarray(eps, dtype=np.float64)
    n = eps.shape[0]
    e_old = np.ascontiguousarray(np.asarray(state["strain"], dtype=np.float64).reshape(n, 6))
    s_old = np.ascontiguousarray(np.asarray(state["stress"], dtype=np.float64).reshape(n, 6))
    p_old = np.ascontiguousarray(np.asarray(state["p"], dtype=np.float64).reshape(n))
    ep_old = np.ascontiguousarray(np.asarray(state["epsp"], dtype=np.float64).reshape(n, 6))
    keep, ptrs, pers = _props(props, n)
    sig, p, epsp, Ct = np.empty((n, 6)), np.empty(n), np.empty((n, 6)), np.empty((n, 6, 6))
    flag, fail = np.empty(n, np.uint8), np.empty(n, np.uint8)
    n_iter, resid = np.empty(n, np.int32), np.empty(n)
    def call(lo, hi):
        lib.dxo_small_strain(ctypes.c_int64(hi - lo), _c(eps, lo), _c(e_old, lo), _c(s_old, lo), _c(p_old, lo),
                             _c(ep_old, lo), ptrs(lo), pers, ctypes.c_int(newton_cap), ctypes.c_double(rtol),
                             _c(sig, lo), _c(p, lo), _c(epsp, lo), _c(Ct, lo), _c(flag, lo), _c(n_iter, lo),
                             _c(resid, lo), _c(fail, lo))

    _run_blocks(n, call)
    del keep
    return {"strain": eps, "stress": sig, "p": p, "epsp": epsp, "Ct": Ct, "flag": flag, "n_iter": n_iter,
            "resid": resid, "fail": fail}


def fefp(F, state, props, newton_cap=25, rtol=1e-12):
    lib = load()
    F = np.ascontiguousarray(F, dtype=np.float64)
    n = F.shape[0]
    Fo = np.ascontiguousarray(np.asarray(state["F"], dtype=np.float64).reshape(n, 9))
    p_old = np.ascontiguousarray(np.asarray(state["p"], dtype=np.float64).reshape(n))
    beo = np.ascontiguousarray(np.asarray(state["be_bar"], dtype=np.float64).reshape(n, 6))
    keep, ptrs, pers = _props(props, n)
    P, p, be, Ct = np.empty((n, 9)), np.empty(n), np.empty((n, 6)), np.empty((n, 9, 9))
    flag, fail = np.empty(n, np.uint8), np.empty(n, np.uint8)
    n_iter, resid = np.empty(n, np.int32), np.empty(n)
    def call(lo, hi):
        lib.dxo_fefp(ctypes.c_int64(hi - lo), _c(F, lo), _c(Fo, lo), _c(p_old, lo), _c(beo, lo), ptrs(lo), pers,
                     ctypes.c_int(newton_cap), ctypes.c_double(rtol), _c(P, lo), _c(p, lo), _c(be, lo), _c(Ct, lo),
                     _c(flag, lo), _c(n_iter, lo), _c(resid, lo), _c(fail, lo))

    _run_blocks(n, call)
    del keep
    return {"F": F, "PK1": P, "p": p, "be_bar": be, "Ct": Ct, "flag": flag, "n_iter": n_iter, "resid": resid,
            "fail": fail}


def hosford_root(q, a):
    """``q**(-1/a)`` as the Hosford criterion computes it (division-free fixed-count iteration, q in (0.5, 1])."""
    lib = load()
    q = np.ascontiguousarray(q, dtype=np.float64)
    w = np.empty_like(q)
    lib.dxo_hosford_root(ctypes.c_int64(q.size), _c(q, 0), ctypes.c_int(int(a)), _c(w, 0))
    return w


def hosford_eig(s):
    """Eigenvalues ``(n, 3)`` and eigenvectors (columns of ``(n, 3, 3)``) of Mandel deviators ``s (n, 6)`` as the
    Hosford update computes them (non-iterative: isolated root of the characteristic cubic, cross product, one Jacobi
    rotation in the normal plane)."""
    lib = load()
    s = np.ascontiguousarray(s, dtype=np.float64).reshape(-1, 6)
    l, Q = np.empty((s.shape[0], 3)), np.empty((s.shape[0], 3, 3))
    lib.dxo_hosford_eig(ctypes.c_int64(s.shape[0]), _c(s, 0), _c(l, 0), _c(Q, 0))
    return l, Q


def hosford(eps, state, props, newton_cap=25, rtol=1e-12):
    """Small-strain Hosford plasticity (``oracle/c/dxm_oracle_hosford.c``); props: E, nu, sig0, H (scalars or per
    point) and the even integer exponent ``a``."""
    lib = load()
    eps = np.ascontiguousarray(eps, dtype=np.float64)
    n = eps.shape[0]
    e_old = np.ascontiguousarray(np.asarray(state["strain"], dtype=np.float64).reshape(n, 6))
    s_old = np.ascontiguousarray(np.asarray(state["stress"], dtype=np.float64).reshape(n, 6))
    p_old = np.ascontiguousarray(np.asarray(state["p"], dtype=np.float64).reshape(n))
    ep_old = np.ascontiguousarray(np.asarray(state["epsp"], dtype=np.float64).reshape(n, 6))
    a = int(props["a"])
    if a < 2 or a % 2 or a > 64 or a != props["a"]:
        raise ValueError("Hosford exponent: an even integer in [2, 64]")
    # sup sigma_eq / seq_Mises (pure shear), with a safety margin: points below it skip the eigen-decomposition
    bound = float(props.get("bound", (2.0 ** (a - 1) + 1.0) ** (1.0 / a) / np.sqrt(3.0) * (1.0 + 1e-9)))
    keep, ptrs, pers = _props({k: v for k, v in props.items() if k not in ("a", "bound")}, n)
    sig, p, epsp, Ct = np.empty((n, 6)), np.empty(n), np.empty((n, 6)), np.empty((n, 6, 6))
    flag, fail = np.empty(n, np.uint8), np.empty(n, np.uint8)
    n_iter, resid = np.empty(n, np.int32), np.empty(n)

    def call(lo, hi):
        lib.dxo_hosford(ctypes.c_int64(hi - lo), _c(eps, lo), _c(e_old, lo), _c(s_old, lo), _c(p_old, lo),
                        _c(ep_old, lo), ptrs(lo), pers, ctypes.c_int(a), ctypes.c_int(newton_cap), ctypes.c_double(rtol),
                        ctypes.c_double(bound), _c(sig, lo), _c(p, lo), _c(epsp, lo), _c(Ct, lo), _c(flag, lo), _c(n_iter, lo),
                        _c(resid, lo), _c(fail, lo))

    _run_blocks(n, call)
    del keep
    return {"strain": eps, "stress": sig, "p": p, "epsp": epsp, "Ct": Ct, "flag": flag, "n_iter": n_iter,
            "resid": resid, "fail": fail}
