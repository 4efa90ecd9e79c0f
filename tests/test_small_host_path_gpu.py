"""Host arrays of at most 2 048 points take the library's small-batch path (``csrc/dxm_api.cu``: kSmallHostPoints -- a
mapped page-locked staging block read / written by the transposition kernels instead of the copy engines and the
three-stream pipeline).  Same bits as the oracle on either side of the threshold, for whole-handle calls, for ranged
calls into a larger handle and with outputs left out."""

import numpy as np
import pytest

from oracle import fefp as ofe
from oracle import small_strain as ss
from oracle import synth

pytestmark = pytest.mark.gpu

VOCE = dict(E=70e3, nu=0.3, sig0=350.0, sigu=500.0, b=1e3)


def j2(jm, n):
    m = jm.CUDAMaterial(jm.vonMisesIsotropicHardening(
        elasticity=jm.LinearElasticIsotropic(E=VOCE["E"], nu=VOCE["nu"]),
        yield_stress=jm.VoceHardening(sig0=VOCE["sig0"], sigu=VOCE["sigu"], b=VOCE["b"])))
    m.set_data_manager(n)
    return m


@pytest.mark.parametrize("n", [1, 2, 17, 2047, 2048, 2049, 2050, 4100])
def test_whole_handle_calls_either_side_of_the_threshold(jm, n):
    m = j2(jm, n)
    st = ss.zero_state(n)
    for k in (1, 2, 3):
        eps = synth.strain(n, 3, 1.25e-2, k, 3)
        flux, isv, Ct = m.integrate(eps)
        ref = ss.integrate(eps, st, VOCE)
        assert np.array_equal(flux, ref["stress"]) and np.array_equal(Ct, ref["Ct"])
        assert np.array_equal(isv[:, 0], ref["p"]) and np.array_equal(isv[:, 1:], ref["epsp"])
        assert m.last_stats.n_plastic == int(ref["flag"].sum()) and m.last_stats.n_fail == 0
        m.data_manager.update()
        st = ss.advance(ref)
    assert ref["flag"].any() or n < 3


def test_small_and_large_windows_of_one_handle_agree(jm):
    """Ranged calls: windows below the threshold (small path) and above it (pipeline) over the same handle and step."""
    n = 9000
    m = j2(jm, n)
    eps = synth.strain(n, 5, 1.5e-2, 1, 1)
    ref = ss.integrate(eps, ss.zero_state(n), VOCE)
    flux, isv, Ct = np.empty((n, 6)), np.empty((n, 7)), np.empty((n, 36))
    plastic = 0
    for lo, hi in ((0, 2), (2, 2050), (2050, 6000), (6000, 6002), (6002, 8050), (8050, 9000)):
        s = m.integrate_range_into(lo, hi - lo, eps[lo:hi], flux[lo:hi], isv[lo:hi], Ct[lo:hi])
        plastic += s.n_plastic
    assert np.array_equal(flux, ref["stress"]) and np.array_equal(Ct.reshape(n, 6, 6), ref["Ct"])
    assert np.array_equal(isv[:, 0], ref["p"]) and np.array_equal(isv[:, 1:], ref["epsp"])
    assert plastic == int(ref["flag"].sum())
    # outputs left out (state only), then fetched from the device state
    m2 = j2(jm, n)
    for lo, hi in ((0, 1000), (1000, 9000)):
        m2.integrate_range_into(lo, hi - lo, eps[lo:hi], None, None, None)
    assert np.array_equal(m2.get_final_state_dict()["stress"].reshape(n, 6), ref["stress"])


def test_finite_strain_small_batch(jm):
    props = dict(E=70e3, nu=0.3, sig0=500.0, sigu=750.0, b=1000.0)
    n = 300
    m = jm.CUDAMaterial(jm.FeFpJ2Plasticity(elasticity=jm.LinearElasticIsotropic(E=props["E"], nu=props["nu"]),
                                            yield_stress=jm.VoceHardening(sig0=props["sig0"], sigu=props["sigu"], b=props["b"])))
    m.set_data_manager(n)
    st = ofe.virgin_state(n)
    for k in (1, 2):
        F = synth.defgrad(n, 1, 3e-2, k, 2)
        flux, isv, Ct = m.integrate(F)
        ref = ofe.integrate(F, st, props)
        assert np.array_equal(flux, ref["PK1"]) and np.array_equal(Ct.reshape(n, 81), ref["Ct"].reshape(n, 81))
        m.data_manager.update()
        st = ofe.advance(ref)
    assert ref["flag"].any()
