"""``yield_stress`` callables, as the reference's scripts pass them to jaxmat behaviours (``tests/test_FeFp_jax.py:14-19``:
``jm.FeFpJ2Plasticity(elasticity=elastic_model, yield_stress=yield_stress)`` with a Python function; old demo
``_plane_stress_elastoplasticity.py:38-44``): the host side recognises the kernels' hardening family behind the callable
(``behaviors.identify_hardening``) and refuses everything else -- nothing is approximated silently."""
import numpy as np
import pytest

import dolfinx_materials_b200 as jm
from dolfinx_materials_b200.behaviors import identify_hardening

EL = jm.LinearElasticIsotropic(E=70e3, nu=0.3)


def test_reference_test_script_callable_is_recognised_exactly():
    # tests/test_FeFp_jax.py:7-19, with numpy standing in for jax.numpy
    sig0, b, sigu = 500.0, 1000, 750.0

    def yield_stress(p):
        return sig0 + (sigu - sig0) * (1 - np.exp(-b * p))

    behavior = jm.FeFpJ2Plasticity(elasticity=EL, yield_stress=yield_stress)
    assert behavior.yield_stress == jm.VoceHardening(sig0=500.0, sigu=750.0, b=1000.0, H=0.0)
    assert behavior.properties() == {"E": 70e3, "nu": 0.3, "sig0": 500.0, "sigu": 750.0, "b": 1000.0, "H": 0.0}
    # the old plane-stress demo's law (E = 70e3, sig0 = 350, sigu = 500, b = 1e3) behind vonMisesIsotropicHardening
    beh = jm.vonMisesIsotropicHardening(elasticity=EL, yield_stress=lambda p: 350.0 + (500.0 - 350.0) * (1 - np.exp(-1e3 * p)))
    assert beh.yield_stress == jm.VoceHardening(sig0=350.0, sigu=500.0, b=1e3) and beh.kind == 2  # DXM_J2_VOCE
    gih = jm.GeneralIsotropicHardening(elasticity=EL, yield_stress=lambda p: 200.0 + 10.0 * p)
    assert gih.yield_stress == jm.LinearHardening(sig0=200.0, H=10.0)


@pytest.mark.parametrize("law", [
    jm.LinearHardening(sig0=250.0, H=5e3),
    jm.LinearHardening(sig0=250.0, H=0.0),
    jm.LinearHardening(sig0=250.0, H=1e-6),  # tests/mfront/test_elastoplasticity.py:21-25
    jm.VoceHardening(sig0=200.0, sigu=300.0, b=10.0),  # demos/multimaterials/multimaterials.py:253-257
    jm.VoceHardening(sig0=400.0, sigu=650.0, b=40.0, H=1500.0),
    jm.VoceHardening(sig0=123.456, sigu=154.706, b=0.37, H=77.7),
    jm.VoceHardening(sig0=350.0, sigu=500.0, b=1e5),
    jm.VoceHardening(sig0=500.0, sigu=420.0, b=25.0, H=300.0),  # softening then hardening
])
def test_family_members_are_recovered(law):
    H, dsu, b = law.H, getattr(law, "sigu", law.sig0) - law.sig0, getattr(law, "b", 0.0)
    got = identify_hardening(lambda p: (law.sig0 + H * p) + dsu * (1.0 - np.exp(-b * p)))
    assert got == law  # the callable's own constants, to the last digit


def test_same_law_written_differently_and_irrational_constants():
    # saturation form sigu - (sigu - sig0) exp(-b p): a few ulp away from the canonical form, same constants recovered
    got = identify_hardening(lambda p: 500.0 - (500.0 - 350.0) * np.exp(-1e3 * p))
    assert got == jm.VoceHardening(sig0=350.0, sigu=500.0, b=1e3)
    # constants with no short decimal form: recovered to the precision the samples carry
    s0, su, bb = 100 * np.pi, 100 * np.pi + 50 * np.e, 1e3 / 7
    got = identify_hardening(lambda p: s0 + (su - s0) * (1 - np.exp(-bb * p)))
    assert got.sig0 == s0 and abs(got.sigu - su) < 1e-9 and abs(got.b - bb) < 1e-8 and abs(got.H) < 1e-9
    p = np.geomspace(1e-9, 5.0, 300)
    mine = (got.sig0 + got.H * p) + (got.sigu - got.sig0) * (1 - np.exp(-got.b * p))
    assert np.abs(mine - (s0 + (su - s0) * (1 - np.exp(-bb * p)))).max() < 1e-11 * su


@pytest.mark.parametrize("fn", [
    lambda p: 1.0 + p ** 0.3,                       # power law
    lambda p: 300.0 * (1 + p / 0.01) ** 0.2,        # Swift
    lambda p: 300.0 + 100 * (1 - np.exp(-50 * p)) + 80 * (1 - np.exp(-2e3 * p)),  # two saturation terms
    lambda p: 300.0 if p < 0.01 else 320.0,          # discontinuous
    lambda p: np.nan,
    lambda p: p.undefined_attribute,                 # cannot be probed with floats
])
def test_other_callables_are_refused_not_approximated(fn):
    assert identify_hardening(fn) is None
    for make in (lambda: jm.vonMisesIsotropicHardening(elasticity=EL, yield_stress=fn),
                 lambda: jm.FeFpJ2Plasticity(elasticity=EL, yield_stress=fn),
                 lambda: jm.GeneralIsotropicHardening(elasticity=EL, yield_stress=fn)):
        with pytest.raises(TypeError, match="yield_stress"):
            make()


def test_tabulated_stays_explicit_and_small_strain_only():
    tab = jm.TabulatedHardening.from_callable(lambda p: 1.0 + p ** 0.3, p_max=0.2, n=16)
    assert jm.vonMisesIsotropicHardening(elasticity=EL, yield_stress=tab).kind == 4  # DXM_J2_TABLE
    with pytest.raises(TypeError):
        jm.FeFpJ2Plasticity(elasticity=EL, yield_stress=tab)
