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
    m.enable_timing(1)  # kernel_ms events are off by default below 262144 points
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


def test_partial_state_write_keeps_gradients_written_into_s1(jm):
    """After update() s1 aliases s0; a resident caller may already have written this step's gradients into s1's
    gradient buffer (dxm_device_ptr / GradientEvaluator) when a partial set_state on generation 1 materialises the
    alias: that must copy flux and internal state only, not the previous step's gradients over the new ones."""
    from dolfinx_materials_b200 import _lib
    from dolfinx_materials_b200._lib import check

    n = 4096
    m = voce(jm, n)
    m.synth_gradients(0, 1.25e-2, 1, 2)
    m.integrate_resident()
    m.data_manager.update()  # s1 now aliases s0
    m.synth_gradients(0, 1.25e-2, 2, 2)  # this step's gradients, written into s1's gradient buffer
    want = synth.strain(n, 0, 1.25e-2, 2, 2)
    p = np.full((n, 1), 0.125)
    check(_lib.load().dxm_set_state(m._h, 1, b"p", p.ctypes.data_as(ctypes.c_void_p), _lib.MEM_HOST), "dxm_set_state")
    assert np.array_equal(m.device_view("strain").T.cpu().numpy(), want)
    assert np.array_equal(m.device_view("p").T.cpu().numpy(), p)
    # ... while the flux rows did come over from s0
    assert np.array_equal(m.device_view("stress", gen=1).cpu().numpy(), m.device_view("stress", gen=0).cpu().numpy())


def test_returned_arrays_outlive_the_material(jm):
    """integrate() hands out views of page-locked buffers: they keep the allocation alive (no dangling views when the
    data manager is re-created or the material is dropped)."""
    n = 3000
    m = voce(jm, n)
    eps = synth.strain(n, 2, 1.25e-2, 1, 1)
    flux, isv, ct = m.integrate(eps)
    ref = ss.integrate(eps, ss.zero_state(n), VOCE)
    part = ct[100:200]  # a slice of a view
    m.set_data_manager(n)  # drops the material's references to its output buffers
    junk = [np.empty((n, 36)) for _ in range(4)]  # give a freed block a chance to be reused
    del m
    gc.collect()
    assert np.array_equal(flux, ref["stress"]) and np.array_equal(part.reshape(100, 36), ref["Ct"].reshape(n, 36)[100:200])
    del junk


def test_page_lock_registry(jm):
    from dolfinx_materials_b200 import _lib
    from dolfinx_materials_b200.material import pin_array

    a = np.zeros(1 << 16)
    u1 = pin_array(a)
    u2 = pin_array(a)  # same range again: reference-counted
    with pytest.raises(_lib.DxmError, match="different size"):
        pin_array(a[: 1 << 12])  # same address, other size: reported, not swallowed
    u1()
    u2()
    b = a[8:]
    u3 = pin_array(a)
    with pytest.raises(_lib.DxmError, match="overlaps"):
        pin_array(b)  # overlaps a live registration at another address
    u3()


def test_singular_finite_strain_state_is_flagged(jm):
    from oracle import fefp

    n = 64
    m = jm.CUDAMaterial(jm.FeFpJ2Plasticity(elasticity=jm.LinearElasticIsotropic(E=70e3, nu=0.3),
                                            yield_stress=jm.VoceHardening(sig0=500.0, sigu=750.0, b=1000.0)), warn_on_failure=False)
    m.set_data_manager(n)
    be = np.tile([1.0, 1, 1, 0, 0, 0], (n, 1))
    be[5] = 0.0
    m.set_initial_state_dict({"be_bar": be})
    F = synth.defgrad(n, 0, 3e-2, 1, 1)
    m.enable_diagnostics()
    P, isv, Ct = m.integrate(F)
    st = fefp.virgin_state(n)
    st["be_bar"] = be
    ref = fefp.integrate(F, st, dict(E=70e3, nu=0.3, sig0=500.0, sigu=750.0, b=1000.0))
    assert m.last_stats.n_fail == 1 and m.diagnostics()[3].tolist() == ref["fail"].tolist()
    ok = np.arange(n) != 5
    assert np.array_equal(P[ok], ref["PK1"][ok]) and np.array_equal(Ct[ok], ref["Ct"][ok])
