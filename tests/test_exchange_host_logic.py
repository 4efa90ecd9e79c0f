"""Host-side logic of QuadratureExchange without a GPU: the chunk bookkeeping, buffer rotation and helper-thread
ordering of the pipelined cell-subset path, and the host-pool gather / scatter it uses, driven with a stand-in material
(a test double that implements the few methods the exchange calls with plain numpy -- NOT a CPU fallback of the
product: the product's materials need the CUDA library and a B200)."""
import numpy as np
import pytest



from qmap_standin import StandInMaterial


@pytest.fixture
def no_pinning(monkeypatch):
    """PinnedArray needs the CUDA runtime; the logic under test does not care where the staging arrays live."""
    import dolfinx_materials_b200.exchange as ex

    class Plain:
        def __init__(self, shape):
            self.array = np.zeros(shape)

    monkeypatch.setattr(ex, "PinnedArray", Plain)
    return ex


@pytest.mark.parametrize("pipeline_points,expect_chunks", [(1 << 19, 0), (1500, 9), (100, 127)])
def test_subset_exchange_bookkeeping(jm, no_pinning, monkeypatch, pipeline_points, expect_chunks):
    ex = no_pinning
    monkeypatch.setattr(ex.QuadratureExchange, "PIPELINE_POINTS", pipeline_points)
    ncell, nqp = 5000, 4
    ntot = ncell * nqp
    rng = np.random.default_rng(0)
    cells = np.sort(rng.choice(ncell, 3041, replace=False))  # odd count -> ragged last chunk
    dofs = (nqp * cells[:, None] + np.arange(nqp)[None, :]).ravel()
    mat = StandInMaterial()
    grad = rng.standard_normal(ntot * 6)
    flux, jac = np.full(ntot * 6, 7.0), np.full(ntot * 36, 7.0)
    isv = {"p": np.zeros(ntot), "epsp": np.zeros(ntot * 6)}
    x = ex.QuadratureExchange(mat, ncell, nqp, {"strain": grad}, {"stress": flux}, isv, jac, cells=cells, pin=False)
    assert (len(x._chunks) if x._chunks else 0) == expect_chunks
    for step in range(2):
        grad[:] = rng.standard_normal(ntot * 6)
        p_old = mat.s0["p"].copy()
        mat.calls.clear()
        stats = x.update()
        g = grad.reshape(-1, 6)[dofs]
        want_flux, want_jac = np.full((ntot, 6), 7.0), np.full((ntot, 36), 7.0)
        if step:
            want_flux, want_jac = prev_flux.copy(), prev_jac.copy()
        want_flux[dofs] = 2.0 * g + p_old
        want_jac[dofs] = (g[:, :, None] * np.arange(1.0, 7.0)[None, None, :]).reshape(-1, 36)
        assert np.array_equal(flux.reshape(-1, 6), want_flux) and np.array_equal(jac.reshape(-1, 36), want_jac)
        assert stats.n_points == len(dofs) and stats.n_plastic == int((g[:, 0] > 0).sum())
        assert stats.max_residual == np.abs(g).max()
        if expect_chunks:
            # every point exactly once, in order, chunk starts even
            assert [c[0] for c in mat.calls] == ["range"] * expect_chunks
            assert sum(c[2] for c in mat.calls) == len(dofs) and mat.calls[0][1] == 0
            assert all(a[1] + a[2] == b[1] for a, b in zip(mat.calls, mat.calls[1:]))
        else:
            assert mat.calls == [("all", 0, len(dofs))]
        x.advance()
        assert np.array_equal(isv["p"][dofs], (p_old + np.abs(g).sum(axis=1, keepdims=True)).ravel())
        assert np.array_equal(isv["epsp"].reshape(-1, 6)[dofs], -g)
        untouched = np.setdiff1d(np.arange(ntot), dofs)
        assert np.all(isv["p"][untouched] == 0) and np.all(flux.reshape(-1, 6)[untouched] == 7.0)
        prev_flux, prev_jac = flux.reshape(-1, 6).copy(), jac.reshape(-1, 36).copy()
    x.close()


def test_subset_cells_are_validated(jm, no_pinning):
    ex = no_pinning
    mat = StandInMaterial()
    arrs = dict(gradients={"strain": np.zeros(60)}, fluxes={"stress": np.zeros(60)},
                internal_state_variables={"p": np.zeros(10), "epsp": np.zeros(60)}, jacobian_flatten=np.zeros(360))
    with pytest.raises(ValueError):
        ex.QuadratureExchange(mat, 10, 1, cells=np.array([1, 1, 2]), pin=False, **arrs)
    with pytest.raises(ValueError):
        ex.QuadratureExchange(mat, 10, 1, cells=np.array([1, 12]), pin=False, **arrs)


class _FiniteStrainStandIn:
    """records what initialize_state pushes; finite-strain protocol names"""
    gradients, fluxes, internal_state_variables = {"F": 9}, {"PK1": 9}, {"p": 1, "be_bar": 6}
    material_properties = {}
    _n = None

    def set_data_manager(self, n):
        self._n = n

    def set_initial_state_dict(self, state):
        self.pushed = {k: np.array(v, copy=True) for k, v in state.items()}


@pytest.mark.parametrize("subset", [False, True])
def test_uninitialised_be_bar_is_seeded_with_the_identity(jm, no_pinning, subset):
    """jaxmat initialises be_bar to the identity by itself (behavior.init_state) and the reference's finite-strain demo
    relies on it; a zero-filled be_bar Function must not overwrite that with a singular state (PK1 would silently be 0)."""
    ex = no_pinning
    ncell, nqp = 12, 4
    ntot = ncell * nqp
    cells = np.array([1, 4, 5, 9]) if subset else None
    arrs = dict(gradients={"F": np.tile([1, 1, 1, 0, 0, 0, 0, 0, 0.0], ntot)}, fluxes={"PK1": np.zeros(ntot * 9)},
                internal_state_variables={"p": np.zeros(ntot), "be_bar": np.zeros(ntot * 6)}, jacobian_flatten=np.zeros(ntot * 81))
    mat = _FiniteStrainStandIn()
    x = ex.QuadratureExchange(mat, ncell, nqp, cells=cells, pin=False, **arrs)
    x.initialize_state()
    ident = np.array([1, 1, 1, 0, 0, 0.0])
    assert np.array_equal(mat.pushed["be_bar"], np.tile(ident, (x.n, 1)))
    be = arrs["internal_state_variables"]["be_bar"].reshape(-1, 6)
    if subset:
        assert np.array_equal(be[x.dofs], np.tile(ident, (x.n, 1)))  # the Function agrees with the device state
        assert not be[np.setdiff1d(np.arange(ntot), x.dofs)].any()  # other regions' points untouched
    else:
        assert np.array_equal(be, np.tile(ident, (ntot, 1)))
    # a be_bar the caller did initialise is pushed as it is
    be[:] = [2.0, 0.5, 1.0, 0.1, 0.0, 0.0]
    x.initialize_state()
    assert np.array_equal(mat.pushed["be_bar"], np.tile([2.0, 0.5, 1.0, 0.1, 0.0, 0.0], (x.n, 1)))
