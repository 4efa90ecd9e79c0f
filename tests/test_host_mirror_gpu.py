"""Packed-tangent hand-off on the host path (``DXM_HOST_MIRROR``): the device sends the 21 unique entries of the
symmetric tangent and host threads mirror them into the reference's ``(n, 36)`` array (``quadrature_map.py:334``).
Both settings must return the same bits as the oracle, for one chunk, several chunks (ring reuse) and ragged tails,
into library-owned pinned outputs and into a caller's plain (pageable) arrays."""

import numpy as np
import pytest

from oracle import small_strain as ss
from oracle import synth

pytestmark = pytest.mark.gpu

VOCE = dict(E=70e3, nu=0.3, sig0=350.0, sigu=500.0, b=1e3)


def material(jm, n):
    beh = jm.vonMisesIsotropicHardening(
        elasticity=jm.LinearElasticIsotropic(E=VOCE["E"], nu=VOCE["nu"]),
        yield_stress=jm.VoceHardening(sig0=VOCE["sig0"], sigu=VOCE["sigu"], b=VOCE["b"]),
    )
    m = jm.CUDAMaterial(beh)
    m.set_data_manager(n)
    return m


@pytest.mark.parametrize("mirror", ["0", "1"])
@pytest.mark.parametrize("n", [1, 4097, 300_001, 2_200_003])
def test_host_path_bit_exact_with_and_without_the_mirror(jm, monkeypatch, mirror, n):
    monkeypatch.setenv("DXM_HOST_MIRROR", mirror)
    m = material(jm, n)
    st = ss.zero_state(n)
    for k in (1, 2):
        eps = synth.strain(n, 3, 1.25e-2, k, 2)
        flux, isv, Ct = m.integrate(eps)
        ref = ss.integrate(eps, st, VOCE)
        assert np.array_equal(flux, ref["stress"])
        assert np.array_equal(isv[:, 0], ref["p"]) and np.array_equal(isv[:, 1:], ref["epsp"])
        assert np.array_equal(Ct, ref["Ct"])
        m.data_manager.update()
        st = ss.advance(ref)
    assert 0 < ref["flag"].sum() < n or n == 1


@pytest.mark.parametrize("mirror", ["0", "1"])
def test_mirror_into_pageable_caller_arrays_and_partial_outputs(jm, monkeypatch, mirror):
    monkeypatch.setenv("DXM_HOST_MIRROR", mirror)
    n = 70_001
    m = material(jm, n)
    eps = synth.strain(n, 5, 1.25e-2, 1, 1)
    ref = ss.integrate(eps, ss.zero_state(n), VOCE)
    ct = np.full((n, 36), np.nan)
    m.integrate_into(eps, ct_out=ct)  # tangent only
    assert np.array_equal(ct, ref["Ct"].reshape(n, 36))
    flux = np.empty((n, 6))
    m.integrate_into(eps, flux_out=flux)  # no tangent: the mirror stage is not entered
    assert np.array_equal(flux, ref["stress"])
    # an odd (8-byte aligned only) destination takes the plain-store branch of the mirror
    buf = np.full(n * 36 + 1, np.nan)
    off = 1 if buf.ctypes.data % 16 == 0 else 0
    m.integrate_into(eps, ct_out=buf[off: off + n * 36])
    assert np.array_equal(buf[off: off + n * 36].reshape(n, 36), ref["Ct"].reshape(n, 36))
