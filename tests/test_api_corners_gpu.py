"""C-ABI corners through ctypes: device-AoS in/out (DXM_MEM_DEVICE), properties from device memory, error
returns (never exceptions across the ABI), caller-supplied stream, several handles on one device, and the
DLPack view keeping its handle alive."""
import ctypes
import gc

import numpy as np
import pytest

from oracle import small_strain as ss
from oracle import synth

pytestmark = pytest.mark.gpu
VOCE = dict(E=70e3, nu=0.3, sig0=350.0, sigu=500.0, b=1e3)


def voce(jm, n):
    m = jm.CUDAMaterial(jm.vonMisesIsotropicHardening(
        elasticity=jm.LinearElasticIsotropic(E=70e3, nu=0.3), yield_stress=jm.VoceHardening(sig0=350.0, sigu=500.0, b=1e3)))
    m.set_data_manager(n)
    return m


def test_device_aos_in_and_out(jm):
    import torch

    from dolfinx_materials_b200 import _lib

    lib = _lib.load()
    n = 70001
    m = voce(jm, n)
    eps = synth.strain(n, 4, 1.25e-2, 1, 1)
    g = torch.from_numpy(eps).cuda()
    flux = torch.empty((n, 6), dtype=torch.float64, device="cuda")
    isv = torch.empty((n, 7), dtype=torch.float64, device="cuda")
    ct = torch.empty((n, 36), dtype=torch.float64, device="cuda")
    stats = _lib.Stats()
    rc = lib.dxm_integrate(m._h, ctypes.c_void_p(g.data_ptr()), _lib.MEM_DEVICE, 0.0, ctypes.c_void_p(flux.data_ptr()),
                           ctypes.c_void_p(isv.data_ptr()), ctypes.c_void_p(ct.data_ptr()), _lib.MEM_DEVICE,
                           ctypes.byref(stats))
    assert rc == 0
    torch.cuda.synchronize()
    ref = ss.integrate(eps, ss.zero_state(n), VOCE)
    assert np.array_equal(flux.cpu().numpy(), ref["stress"])
    assert np.array_equal(isv.cpu().numpy()[:, 0], ref["p"]) and np.array_equal(isv.cpu().numpy()[:, 1:], ref["epsp"])
    assert np.array_equal(ct.cpu().numpy().reshape(n, 6, 6), ref["Ct"])
    assert stats.n_plastic == int(ref["flag"].sum()) and stats.n_points == n and stats.kernel_ms > 0


def test_property_from_device_memory_and_errors(jm):
    import torch

    from dolfinx_materials_b200 import _lib

    lib = _lib.load()
    n = 1000
    m = voce(jm, n)
    sig0 = torch.full((n,), 350.0, dtype=torch.float64, device="cuda")
    sig0[::2] = 1e9  # every other point stays elastic
    assert lib.dxm_set_property(m._h, b"sig0", ctypes.c_void_p(sig0.data_ptr()), n, _lib.MEM_DEVICE) == 0
    eps = synth.strain(n, 4, 1.25e-2, 1, 1)
    m.enable_diagnostics()
    m.integrate(eps)
    props = dict(VOCE, sig0=sig0.cpu().numpy(), sigu=np.where(np.arange(n) % 2 == 0, 500.0, 500.0))
    ref = ss.integrate(eps, ss.zero_state(n), props)
    assert np.array_equal(m.diagnostics()[0], ref["flag"]) and ref["flag"][::2].sum() == 0
    # errors come back as negative codes with a message, never as a crash
    assert lib.dxm_set_property(m._h, b"bogus", ctypes.c_void_p(sig0.data_ptr()), 1, _lib.MEM_DEVICE) < 0
    assert b"bogus" in lib.dxm_last_error()
    assert lib.dxm_set_property(m._h, b"E", ctypes.c_void_p(sig0.data_ptr()), n - 1, _lib.MEM_DEVICE) < 0
    buf = np.zeros((n, 6))
    assert lib.dxm_get_state(m._h, 1, b"nope", buf.ctypes.data_as(ctypes.c_void_p), _lib.MEM_HOST) < 0
    assert lib.dxm_get_state(m._h, 3, b"stress", buf.ctypes.data_as(ctypes.c_void_p), _lib.MEM_HOST) < 0
    h = ctypes.c_void_p()
    assert lib.dxm_create(99, 0, 10, ctypes.byref(h)) < 0 and lib.dxm_create(2, 0, 0, ctypes.byref(h)) < 0
    assert lib.dxm_create(2, 64, 10, ctypes.byref(h)) < 0 and b"no CPU fallback" in lib.dxm_last_error()
    with pytest.raises(KeyError):
        m._get_state(0, "nope")
    with pytest.raises(ValueError):
        m.integrate(eps[:-1])
    fresh = jm.CUDAMaterial(jm.ElasticBehavior(elasticity=jm.LinearElasticIsotropic(E=1.0, nu=0.2)))
    with pytest.raises(RuntimeError):
        fresh.integrate(eps)  # set_data_manager not called


def test_caller_stream_and_two_handles(jm):
    import torch

    n = 200000
    a, b = voce(jm, n), voce(jm, n // 2)
    s = torch.cuda.Stream()
    a.set_stream(s.cuda_stream)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.synth_gradients(0, 1.25e-2, 1, 1)
    b.synth_gradients(0, 1.25e-2, 1, 1)
    with torch.cuda.stream(s):
        ev0.record(s)
        a.integrate_resident(wait=False)
        ev1.record(s)
    sb = b.integrate_resident()
    s.synchronize()
    sa = a.fetch_stats()
    assert ev0.elapsed_time(ev1) > 0  # the kernel really ran on the caller's stream
    ra = ss.integrate(synth.strain(n, 0, 1.25e-2, 1, 1), ss.zero_state(n), VOCE)
    assert sa.n_plastic == int(ra["flag"].sum()) and sb.n_plastic == int(ra["flag"][: n // 2].sum())
    assert np.array_equal(b.device_view("stress").cpu().numpy().T, ra["stress"][: n // 2])


def test_dlpack_view_outlives_material(jm):
    n = 5000
    m = voce(jm, n)
    m.synth_gradients(0, 1.25e-2, 1, 1)
    m.integrate_resident()
    view = m.device_view("stress")
    expected = view.clone()
    del m
    gc.collect()
    assert bool((view == expected).all())  # buffers are freed only when the last view dies
    del view
    gc.collect()
