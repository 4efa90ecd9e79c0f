"""The FE kernels' per-cell / per-row routines (``csrc/dxm_fe_gradient.cuh``: ``fe_gradient_cell``;
``csrc/dxm_fe_forms.cuh``: ``fe_form_point_geometry`` / ``fe_form_row``, all ``__host__ __device__``) executed on the CPU
and compared bit for bit with the oracles -- the code the GPU runs per cell / per element-matrix row, checked where no
GPU is available.  (GPU parity tests proper: ``tests/test_fe_gradient_gpu.py``, ``tests/test_fe_forms_gpu.py``.)"""

import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import fe_forms as ff
from oracle import fe_gradient as fg
from oracle import fefp
from oracle import small_strain as ss

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "fe_host_check.cu")
LIB = os.path.join(HERE, "_build", "libfe_host_check.so")
SYM6 = np.array([min(c // 6, c % 6) * 6 - (min(c // 6, c % 6) * (min(c // 6, c % 6) - 1)) // 2 + abs(c // 6 - c % 6)
                 for c in range(36)])  # sym6_packed (include/dxm.h)


@pytest.fixture(scope="module")
def host():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    deps = [SRC] + [os.path.join(HERE, "..", "dolfinx_materials_b200", "csrc", f)
                    for f in ("dxm_fe_forms.cuh", "dxm_fe_gradient.cuh", "dxm_canon.cuh")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-fmad=false",
                        "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-o", LIB, SRC], check=True)
    return ctypes.CDLL(LIB)


def c(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def field(nodes, amp):
    x, y, z = nodes.T
    return amp * np.stack([x * y + 0.5 * z * z + 0.3 * x, -2 * y * z + 0.3 * x * x - 0.2 * y, 0.7 * x * z - 0.4 * y * y + 0.1 * z], axis=1)


def run_gradient(lib, coords, gd, ud, u, dphi, kind, tdim, generic=0):
    coords, dphi, u = (np.ascontiguousarray(a, dtype=np.float64) for a in (coords, dphi, u))
    gd, ud = np.ascontiguousarray(gd, dtype=np.int32), np.ascontiguousarray(ud, dtype=np.int32)
    nc, nd = ud.shape
    nqp = dphi.shape[0]
    n = nc * nqp
    ld = (n + 63) & ~63
    out = np.zeros((6 if kind == 0 else 9, ld))
    rc = lib.fe_gradient_host(ctypes.c_int(tdim), ctypes.c_int64(nc), ctypes.c_int(nd), ctypes.c_int(nqp), ctypes.c_int(kind),
                              c(coords), c(gd), c(ud), c(u), c(dphi), c(out), ctypes.c_int64(ld), ctypes.c_int(generic))
    assert rc == 0
    return np.ascontiguousarray(out[:, :n].T)


def run_forms(lib, coords, gd, ud, dphi, w, flux, ct, kind, tdim, want_mat=1, generic=0):
    coords, dphi, w = (np.ascontiguousarray(a, dtype=np.float64) for a in (coords, dphi, w))
    gd, ud = np.ascontiguousarray(gd, dtype=np.int32), np.ascontiguousarray(ud, dtype=np.int32)
    nc, nd = ud.shape
    nqp = dphi.shape[0]
    n = nc * nqp
    ld = (n + 63) & ~63
    nf = 6 if kind == 0 else 9
    fl = np.zeros((nf, ld))
    fl[:, :n] = flux.T
    full = ct.reshape(n, nf * nf)
    rows = np.array([int(np.flatnonzero(SYM6 == k)[0]) for k in range(21)]) if kind == 0 else np.arange(81)
    cts = np.zeros((len(rows), ld))
    cts[:, :n] = full[:, rows].T  # resident layout: packed symmetric rows for the small-strain tangent
    ndof = nd * tdim
    fe, ke = np.zeros((nc, ndof)), np.zeros((nc, ndof, ndof))
    rc = lib.fe_forms_host(ctypes.c_int(tdim), ctypes.c_int64(nc), ctypes.c_int(nd), ctypes.c_int(nqp), ctypes.c_int(kind),
                           c(coords), c(gd), c(ud), c(dphi), c(w), c(fl), c(cts), ctypes.c_int64(ld), ctypes.c_int(want_mat),
                           c(fe), c(ke), ctypes.c_int(generic))
    assert rc == 0
    return fe, ke


def tri_mesh(nx):
    xs = np.linspace(0, 1, nx + 1)
    X, Y = np.meshgrid(xs, xs ** 1.3, indexing="ij")
    c2 = np.stack([X.ravel(), Y.ravel(), np.zeros(X.size)], axis=1)
    nid = lambda i, j: i * (nx + 1) + j  # noqa: E731
    tri = np.array([[nid(i, j), nid(i + 1, j), nid(i + 1, j + 1)] for i in range(nx) for j in range(nx)]
                   + [[nid(i, j), nid(i + 1, j + 1), nid(i, j + 1)] for i in range(nx) for j in range(nx)], dtype=np.int32)
    dphi = np.broadcast_to(np.array([[-1.0, -1.0], [1, 0], [0, 1]]), (1, 3, 2)).copy()
    return c2, tri, dphi


@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("kind", [0, 1])
def test_gradient_cell_routine_equals_oracle(host, order, kind):
    coords, gd, ud, nodes = fg.box_tets(5, 4, 3, order)
    dphi = fg.tet_dphi(fg.TET_QP_DEG1 if order == 1 else fg.TET_QP_DEG2, order)
    u = field(nodes, 0.02).ravel()
    ref = fg.evaluate(coords, gd, ud, u, dphi, kind, 3)
    for generic in (0, 1):  # compile-time P1 / P2 instantiation and the run-time-nd loop
        assert np.array_equal(run_gradient(host, coords, gd, ud, u, dphi, kind, 3, generic), ref), generic


def test_gradient_cell_routine_2d(host):
    c2, tri, dphi = tri_mesh(12)
    u2 = (0.01 * np.stack([c2[:, 0] * c2[:, 1], c2[:, 0] ** 2 - c2[:, 1]], axis=1)).ravel()
    for kind in (0, 1):
        ref = fg.evaluate(c2, tri, tri, u2, dphi, kind, 2)
        for generic in (0, 1):
            got = run_gradient(host, c2, tri, tri, u2, dphi, kind, 2, generic)
            assert np.array_equal(got, ref), (kind, generic)
    assert np.count_nonzero(fg.evaluate(c2, tri, tri, u2, dphi, 0, 2)[:, [2, 4, 5]]) == 0  # plane-strain padding


@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("finite", [False, True])
def test_form_row_routine_equals_oracle_after_a_constitutive_update(host, order, finite):
    """u -> gradients -> oracle update -> element vectors / matrices through the kernel's row routine == oracle."""
    coords, gd, ud, nodes = fg.box_tets(4, 3, 3, order)
    qp = fg.TET_QP_DEG1 if order == 1 else fg.TET_QP_DEG2
    w = np.full(len(qp), 1.0 / 6.0 / len(qp))
    dphi = fg.tet_dphi(qp, order)
    n = len(gd) * len(qp)
    kind = 1 if finite else 0
    grads = fg.evaluate(coords, gd, ud, field(nodes, 0.004).ravel(), dphi, kind, 3)
    if finite:
        res = fefp.integrate(grads, fefp.virgin_state(n), dict(E=70e3, nu=0.3, sig0=500.0, sigu=750.0, b=1000.0))
        flux, ct = res["PK1"], res["Ct"]
    else:
        res = ss.integrate(grads, ss.zero_state(n), dict(E=70e3, nu=0.3, sig0=350.0, sigu=500.0, b=1e3))
        flux, ct = res["stress"], res["Ct"]
    assert 0 < res["flag"].sum() < n and res["fail"].sum() == 0  # elastic and plastic tangents in one mesh
    fe_ref, ke_ref = ff.element_forms(coords, gd, ud, dphi, w, flux, ct, kind, 3)
    for generic in (0, 1):
        fe, ke = run_forms(host, coords, gd, ud, dphi, w, flux, ct, kind, 3, 1, generic)
        assert np.array_equal(fe, fe_ref) and np.array_equal(ke, ke_ref), generic
    fe, _ = run_forms(host, coords, gd, ud, dphi, w, flux, ct, kind, 3, 0)
    assert np.array_equal(fe, fe_ref)


def test_form_row_routine_2d(host):
    c2, tri, dphi = tri_mesh(9)
    rng = np.random.default_rng(5)
    n = len(tri)
    w = np.array([0.5])
    for kind, nf in ((0, 6), (1, 9)):
        flux = rng.standard_normal((n, nf))
        ct = rng.standard_normal((n, nf, nf))
        if kind == 0:
            ct = ct + ct.transpose(0, 2, 1)  # the resident small-strain tangent is stored symmetric-packed
        fe_ref, ke_ref = ff.element_forms(c2, tri, tri, dphi, w, flux, ct, kind, 2)
        for generic in (0, 1):
            fe, ke = run_forms(host, c2, tri, tri, dphi, w, flux, ct, kind, 2, 1, generic)
            assert np.array_equal(fe, fe_ref) and np.array_equal(ke, ke_ref), (kind, generic)
