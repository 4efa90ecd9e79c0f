"""The drop-in ``GPUQuadratureMap`` (dolfinx_materials_b200/quadrature_map.py) against the class it replaces.
dolfinx is not installable here, so the base class is the stand-in of tests/qmap_standin.py, which reproduces the
attributes and the update / advance / initialize_state / update_initial_state semantics of the reference's
``QuadratureMap`` on plain arrays; the adapter subclass and the unmodified base must leave the same values in the
same Function arrays.  CPU: stand-in material (host logic only).  GPU: CUDA materials."""
import numpy as np
import pytest

from qmap_standin import StandInMaterial, StandInQuadratureMap

from dolfinx_materials_b200.quadrature_map import gpu_quadrature_map


def drive(qmap, ntot, gdim, gname, gen, steps, init=None):
    """The demos' call sequence: register, optional update_initial_state, first update at the initial gradients,
    then load steps of a few Newton iterations each followed by advance()."""
    current = {"g": gen(0)}
    qmap.register_gradient(gname, lambda: current["g"])
    if init is not None:
        qmap.update_initial_state(*init)
    qmap.update()
    snaps = []
    for step in range(1, steps + 1):
        for scale in (0.7, 1.0):
            current["g"] = gen(0) + scale * (gen(step) - gen(0))
            qmap.update()
        qmap.advance()
        snaps.append({k: f.x.array.copy() for k, f in {**qmap.fluxes, **qmap.internal_state_variables, "Ct": qmap.jacobian_flatten}.items()})
    return snaps


@pytest.fixture
def plain_staging(monkeypatch):
    import dolfinx_materials_b200.exchange as ex

    class Plain:
        def __init__(self, shape):
            self.array = np.zeros(shape)

    monkeypatch.setattr(ex, "PinnedArray", Plain)
    monkeypatch.setattr(ex, "pin_array", lambda a: (lambda: None))


@pytest.mark.parametrize("subset", [False, True])
def test_adapter_equals_base_class_with_a_stand_in_material(jm, plain_staging, subset):
    ncell, nqp = 700, 4
    ntot = ncell * nqp
    cells = np.sort(np.random.default_rng(1).choice(ncell, 431, replace=False)) if subset else None
    rng = np.random.default_rng(0)
    fields = [np.zeros((ntot, 6))] + [rng.standard_normal((ntot, 6)) for _ in range(3)]
    gen = lambda k: fields[k]  # noqa: E731
    GPUQuadratureMap = gpu_quadrature_map(StandInQuadratureMap)
    ref = drive(StandInQuadratureMap(ncell, nqp, StandInMaterial(), cells=cells), ntot, 6, "strain", gen, 3, init=("p", 0.25))
    mat = StandInMaterial()
    q = GPUQuadratureMap(ncell, nqp, mat, cells=cells)
    got = drive(q, ntot, 6, "strain", gen, 3, init=("p", 0.25))
    for a, b in zip(ref, got):
        for k in a:
            assert np.array_equal(a[k], b[k]), k
    assert not any(c[0] == "integrate" for c in mat.calls)  # the reference-protocol integrate() is never used
    assert q.last_stats.n_points == (len(cells) if subset else ncell) * nqp
    q.close()


def test_internal_state_functions_can_follow_every_update(jm, plain_staging):
    """internal_state_every_update=True: the isv Functions hold the same values as the reference's after EVERY update
    (quadrature_map.py:333), not only after advance()."""
    ncell, nqp = 300, 3
    ntot = ncell * nqp
    cells = np.arange(ncell)[::2]
    g = np.random.default_rng(4).standard_normal((ntot, 6))
    base = StandInQuadratureMap(ncell, nqp, StandInMaterial(), cells=cells)
    q = gpu_quadrature_map(StandInQuadratureMap, internal_state_every_update=True)(ncell, nqp, StandInMaterial(), cells=cells)
    lazy = gpu_quadrature_map(StandInQuadratureMap)(ncell, nqp, StandInMaterial(), cells=cells)
    for m in (base, q, lazy):
        m.register_gradient("strain", lambda: g)
        m.update()
    for k in base.internal_state_variables:
        assert np.array_equal(q.internal_state_variables[k].x.array, base.internal_state_variables[k].x.array)
    assert not np.array_equal(lazy.internal_state_variables["epsp"].x.array, base.internal_state_variables["epsp"].x.array)
    lazy.advance()
    base.advance()
    assert np.array_equal(lazy.internal_state_variables["epsp"].x.array, base.internal_state_variables["epsp"].x.array)


def test_adapter_requires_registered_gradients(jm, plain_staging):
    q = gpu_quadrature_map(StandInQuadratureMap)(10, 1, StandInMaterial())
    with pytest.raises(ValueError):
        q.update()


@pytest.mark.gpu
@pytest.mark.parametrize("subset", [False, True])
@pytest.mark.parametrize("kind", ["j2_voce", "fefp", "hosford"])
def test_adapter_equals_base_class_with_cuda_materials(jm, subset, kind):
    from oracle import synth

    el = jm.LinearElasticIsotropic(E=70e3, nu=0.3)

    def material():
        if kind == "fefp":
            return jm.CUDAMaterial(jm.FeFpJ2Plasticity(elasticity=el, yield_stress=jm.VoceHardening(sig0=500.0, sigu=750.0, b=1000.0)))
        if kind == "hosford":
            return jm.CUDAMaterial(jm.GeneralIsotropicHardening(elasticity=el, yield_stress=jm.LinearHardening(sig0=200.0, H=10.0)))
        return jm.CUDAMaterial(jm.vonMisesIsotropicHardening(elasticity=el, yield_stress=jm.VoceHardening(sig0=350.0, sigu=500.0, b=1e3)))

    ncell, nqp = 3000, 4
    ntot = ncell * nqp
    cells = np.sort(np.random.default_rng(2).choice(ncell, 1801, replace=False)) if subset else None
    if kind == "fefp":
        gname, gdim = "F", 9
        g0 = np.tile([1, 1, 1, 0, 0, 0, 0, 0, 0.0], (ntot, 1))
        gen = lambda k: g0 if k == 0 else synth.defgrad(ntot, 1, 3e-2, k, 3)  # noqa: E731
        init = ("be_bar", np.array([1, 1, 1, 0, 0, 0.0]))  # finite_strain_elastoplasticity.py:181
    else:
        gname, gdim = "strain", 6
        gen = lambda k: np.zeros((ntot, 6)) if k == 0 else synth.strain(ntot, 1, 1.25e-2, k, 3)  # noqa: E731
        init = None
    ref = drive(StandInQuadratureMap(ncell, nqp, material(), cells=cells), ntot, gdim, gname, gen, 3, init=init)
    q = gpu_quadrature_map(StandInQuadratureMap)(ncell, nqp, material(), cells=cells)
    got = drive(q, ntot, gdim, gname, gen, 3, init=init)
    for a, b in zip(ref, got):
        for k in a:
            assert np.array_equal(a[k], b[k]), k
    assert q.last_stats.n_fail == 0 and q.last_stats.n_plastic > 0
    q.close()
