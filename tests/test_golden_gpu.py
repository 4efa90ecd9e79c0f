"""The CUDA path against the committed golden fixtures (tests/golden/*.npz, produced by running the reference's
own Material / DataManager code in the build container -- tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
from golden_check import close

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_elastic_vs_reference_run(jm):
    """elastic_reference.npz was computed by the reference's LinearElasticIsotropic through its own
    Material.integrate (python_materials/elasticity.py:21-24, generic.py:176-189)."""
    g = np.load(os.path.join(GOLD, "elastic_reference.npz"))
    n = g["eps1"].shape[0]
    m = jm.CUDAMaterial(jm.ElasticBehavior(elasticity=jm.LinearElasticIsotropic(E=70e3, nu=0.3)))
    m.set_data_manager(n)
    for k in (1, 2):
        flux, isv, Ct = m.integrate(g[f"eps{k}"])
        # the reference evaluates sigma = C @ eps (total form); the kernel accumulates stress increments
        np.testing.assert_allclose(flux, g[f"flux{k}"], rtol=1e-10, atol=1e-10)
        assert np.array_equal(Ct, g[f"Ct{k}"])  # exactly the reference's C
        m.data_manager.update()
        np.testing.assert_allclose(m.get_initial_state_dict()["stress"], g[f"s0_stress_after_update{k}"], rtol=1e-10, atol=1e-10)
    m0 = jm.CUDAMaterial(jm.ElasticBehavior(elasticity=jm.LinearElasticIsotropic(E=70e3, nu=0.0)))
    m0.set_data_manager(1)
    flux, _, _ = m0.integrate(np.array([[1e-3, 0, 0, 0, 0, 0.0]]))
    assert np.allclose(flux, g["flux_nu0"]) and np.allclose(flux[0, :3], 70e3 * np.array([1e-3, 0, 0]))


@pytest.mark.parametrize("name", ["j2_voce_history.npz", "j2_linear_history.npz", "fefp_history.npz", "hosford_history.npz"])
def test_histories_vs_reference_protocol_run(jm, name):
    g = np.load(os.path.join(GOLD, name))
    p = dict(zip([str(k) for k in g["props_keys"]], [float(v) for v in g["props_vals"]]))
    el = jm.LinearElasticIsotropic(E=p["E"], nu=p["nu"])
    if "b" in p:
        hard = jm.VoceHardening(sig0=p["sig0"], sigu=p["sigu"], b=p["b"])
    else:
        hard = jm.LinearHardening(sig0=p["sig0"], H=p["H"])
    finite = name.startswith("fefp")
    if finite:
        beh = jm.FeFpJ2Plasticity(elasticity=el, yield_stress=hard)
    elif name.startswith("hosford"):
        beh = jm.GeneralIsotropicHardening(elasticity=el, yield_stress=hard, equivalent_stress=jm.Hosford(a=int(p["a"])))
    else:
        beh = jm.vonMisesIsotropicHardening(elasticity=el, yield_stress=hard)
    key = "F" if finite else "eps"
    n = g[f"{key}1"].shape[0]
    m = jm.CUDAMaterial(beh)
    m.set_data_manager(n)
    k = 1
    while f"{key}{k}" in g:
        flux, isv, Ct = m.integrate(g[f"{key}{k}"])
        # the fixtures hold the round-1 (un-fused) arithmetic; the kernels' fused canonical arithmetic agrees with it
        # to the north star's rtol 1e-10 (bit-identity with the fused oracle: the other GPU tests)
        close(flux, g[f"flux{k}"], "flux")
        for c0, c1 in ((0, 1), (1, isv.shape[1])):
            close(isv[:, c0:c1], g[f"isv{k}"][:, c0:c1], "isv")
        close(Ct, g[f"Ct{k}"], "Ct")
        m.data_manager.update()
        k += 1
    assert k > 3
