"""GPU-aware replacement for the material-facing half of ``QuadratureMap.update`` / ``advance``
(reference ``dolfinx_materials/quadrature_map.py:297-360``) -- SURVEY.md section 8(f) rank 1.

What the reference does around every ``integrate`` call, on the host, per Newton iteration:
gather ``_get_vals(gradient)[dofs, :]`` (a copy), ``np.concatenate``, three ``np.isnan`` scans over
flux / isv / Ct, then fancy-index scatters ``fun.x.array[dofs] = arr`` for the flux, every internal state
variable and the flattened tangent (``utils.py:136-143``).  Once the update itself runs at HBM speed these
passes dominate.  ``QuadratureExchange`` keeps the same observable result (the same values end up in the same
``x.array`` vectors) and removes them:

* contiguous-range fast path: when the map covers all cells (``cells is None`` in the reference, the common
  case) ``dofs`` is the identity, so the gradient ``x.array`` is handed to the library as is and the device
  DMAs flux and tangent straight into the (page-locked) ``x.array`` of their Functions;
* the NaN scans are replaced by the fail count reduced on the device;
* internal state variables stay on the GPU during Newton iterations and are fetched once, in ``advance()``.

Cell subsets (one map per material region, ``demos/multimaterials/multimaterials.py:265-273``) cannot avoid one
gather into / scatter out of page-locked staging arrays -- the Function arrays span the whole mesh -- but they run
cell-block-wise on the library's host thread pool (``dxm_host_gather_rows`` / ``dxm_host_scatter_rows``) instead
of numpy fancy indexing.  The class works on any objects exposing a flat float64 ``x.array`` (dolfinx
``fem.Function``) or on plain ndarrays, so it runs without dolfinx.
"""

import ctypes
import warnings
import weakref
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import PerformanceWarning, _lib
from .material import PinnedArray, pin_array


def _gather_cells(src_flat, cells64, block, dst):
    """dst[c, :] = src[cells[c], :] with rows of ``block`` doubles (all points of a cell), on the host pool."""
    _lib.check(_lib.load().dxm_host_gather_rows(src_flat.ctypes.data_as(ctypes.c_void_p), cells64.ctypes.data_as(ctypes.c_void_p),
                                                len(cells64), block, dst.ctypes.data_as(ctypes.c_void_p), 0), "dxm_host_gather_rows")


def _scatter_cells(dst_flat, cells64, block, src):
    """dst[cells[c], :] = src[c, :] (``_update_vals`` with ``cells``, ``utils.py:136-143``), on the host pool."""
    _lib.check(_lib.load().dxm_host_scatter_rows(dst_flat.ctypes.data_as(ctypes.c_void_p), cells64.ctypes.data_as(ctypes.c_void_p),
                                                 len(cells64), block, src.ctypes.data_as(ctypes.c_void_p), 0), "dxm_host_scatter_rows")


def _run_all(callbacks):
    while callbacks:
        callbacks.pop()()


def _flat(fun):
    """``fun.x.array`` for Function-like objects, the array itself otherwise."""
    x = getattr(fun, "x", None)
    arr = x.array if x is not None else getattr(fun, "array", fun)
    if not isinstance(arr, np.ndarray) or arr.dtype != np.float64 or not arr.flags.c_contiguous:
        raise TypeError("expected a C-contiguous float64 array (or an object with .x.array)")
    return arr


class QuadratureExchange:
    PIPELINE_POINTS = 1 << 19  # points per pipelined chunk of a cell-subset map (the library's own transfer chunk)

    def __init__(self, material, num_cells, num_qp, gradients, fluxes, internal_state_variables, jacobian_flatten,
                 cells=None, pin=True, strict=True, keep_data_manager=False):
        """``keep_data_manager``: the material already has a data manager of the right size (the reference's
        ``QuadratureMap.__init__`` created it, ``quadrature_map.py:231-233``) whose properties and initial state must
        survive -- do not create a new one."""
        self.material = material
        self.strict = strict
        self.num_qp = int(num_qp)
        if len(material.gradients) != 1 or len(material.fluxes) != 1:
            raise NotImplementedError("single-gradient / single-flux materials only (as the CUDA behaviours are)")
        self.gname, self.gdim = next(iter(material.gradients.items()))
        self.fname, self.fdim = next(iter(material.fluxes.items()))
        self.grad = _flat(gradients[self.gname])
        self.flux = _flat(fluxes[self.fname])
        self.isv = {k: _flat(internal_state_variables[k]) for k in material.internal_state_variables}
        self.jac = _flat(jacobian_flatten)
        ntot = int(num_cells) * self.num_qp
        self.identity = cells is None or (
            len(cells) == num_cells and np.array_equal(np.asarray(cells), np.arange(num_cells))
        )
        if self.identity:
            self.cells = np.arange(num_cells, dtype=np.int32)
            self.dofs = None
            self.n = ntot
        else:
            self.cells = np.asarray(cells, dtype=np.int32)
            if len(np.unique(self.cells)) != len(self.cells) or self.cells.min() < 0 or self.cells.max() >= num_cells:
                raise ValueError("cells must be distinct indices in [0, num_cells)")
            self.cells64 = np.ascontiguousarray(self.cells, dtype=np.int64)
            self.dofs = (np.repeat(self.num_qp * self.cells[:, None], self.num_qp, axis=1)
                         + np.arange(self.num_qp)[None, :]).ravel()
            self.n = len(self.dofs)
        if keep_data_manager:
            if getattr(material, "_n", None) != self.n:
                raise ValueError(f"keep_data_manager: the material's data manager holds {getattr(material, '_n', None)} points, the map {self.n}")
        else:
            material.set_data_manager(self.n)
            for name, prop in material.material_properties.items():
                material.update_material_property(name, np.asarray(prop))
        self._unpin = []
        # the page-locked registrations are released with the exchange even when close() is never called
        self._fin = weakref.finalize(self, _run_all, self._unpin)
        self._stage = None
        if self.identity and pin:
            for a in (self.grad, self.flux, self.jac, *self.isv.values()):
                self._unpin.append(pin_array(a))
        elif not self.identity:
            nisv = sum(material.internal_state_variables.values())
            # large subsets: chunks of whole cells, double-buffered, so that the host gather / scatter of one chunk
            # overlaps the transfers and the update of its neighbours (integrate_range_into)
            per = max(2, (self.PIPELINE_POINTS // self.num_qp) & ~1)  # cells per chunk, even -> even point offsets
            self._chunks = ([(c0, min(per, len(self.cells) - c0)) for c0 in range(0, len(self.cells), per)]
                            if self.n >= 2 * self.PIPELINE_POINTS else None)
            m = per * self.num_qp if self._chunks else self.n
            nbuf = 2 if self._chunks else 1
            self._stage = [(PinnedArray((m, self.gdim)), PinnedArray((m, self.fdim)),
                            PinnedArray((m, self.fdim * self.gdim))) for _ in range(nbuf)]
            self._stage_state = PinnedArray((self.n, max(nisv, self.fdim, 1)))
            self._host = ThreadPoolExecutor(1) if self._chunks else None
        self._initialized = False
        self.last_stats = None

    def close(self):
        _run_all(self._unpin)
        if getattr(self, "_host", None) is not None:
            self._host.shutdown()
            self._host = None

    # ---- quadrature_map.py:281-295 -------------------------------------------------------------------
    def _take(self, flat, dim):
        vals = flat.reshape(-1, max(1, dim))
        return vals if self.identity else vals[self.dofs]

    def initialize_state(self):
        state = {self.gname: self._take(self.grad, self.gdim), self.fname: self._take(self.flux, self.fdim)}
        for k, d in self.material.internal_state_variables.items():
            state[k] = self._take(self.isv[k], d)
        # Finite strain: jaxmat initialises be_bar (isochoric elastic left Cauchy-Green) to the IDENTITY by itself
        # (behavior.init_state, jaxmat.py:35), and the reference's finite-strain demo relies on that when it never calls
        # update_initial_state("be_bar").  A Function that was never initialised holds zeros -- a singular state that
        # would silently give PK1 = 0 -- so an all-zero be_bar is seeded with the identity, in the Function too.
        if "be_bar" in state and not state["be_bar"].any():
            ident = np.array([1.0, 1.0, 1.0, 0.0, 0.0, 0.0])
            vals = self.isv["be_bar"].reshape(-1, 6)
            if self.identity:
                vals[:] = ident
            else:
                vals[self.dofs] = ident
            state["be_bar"] = self._take(self.isv["be_bar"], 6)
        self.material.set_initial_state_dict(state)
        self._initialized = True

    def update_initial_state(self, field_name, value):
        arrs = {self.gname: (self.grad, self.gdim), self.fname: (self.flux, self.fdim)}
        arrs.update({k: (self.isv[k], d) for k, d in self.material.internal_state_variables.items()})
        flat, dim = arrs[field_name]
        vals = flat.reshape(-1, max(1, dim))
        new = np.full((self.n, max(1, dim)), value, dtype=np.float64)
        if self.identity:
            vals[:] = new
        else:
            vals[self.dofs] = new
        self.material.set_initial_state_dict({field_name: new})

    # ---- quadrature_map.py:297-334 ---------------------------------------------------------------------
    def update(self, dt=0, fetch_internal_state=False):
        """The gradient ``x.array`` must hold this iteration's evaluated gradients (what
        ``QuadratureExpression.eval`` leaves there, ``quadrature_function.py:45-51``).  ``fetch_internal_state``: also
        refresh the internal-state Functions now, as the reference does on every update (``quadrature_map.py:333``);
        by default they are refreshed in ``advance()`` only."""
        if not self._initialized:
            self.initialize_state()
        m = self.material
        if self.identity:
            stats = m.integrate_into(self.grad, self.flux, None, self.jac, dt)
        else:
            stats = self._update_subset(dt)
        self.last_stats = stats
        if fetch_internal_state:
            self.fetch_state(flux=False)
        if stats.n_fail:
            # the reference asserts on NaN in flux / isv / Ct (quadrature_map.py:322-324); the device-side fail
            # count covers those (non-finite results) plus local solves that hit their iteration cap
            msg = f"{stats.n_fail} Gauss point(s) failed their constitutive update (max residual {stats.max_residual:.3e})"
            if self.strict:
                raise AssertionError(msg)
            warnings.warn(msg, PerformanceWarning)
        return stats

    def _update_subset(self, dt):
        m, q = self.material, self.num_qp
        gb, fb, cb = q * self.gdim, q * self.fdim, q * self.fdim * self.gdim
        if not self._chunks:
            g, f, c = self._stage[0]
            _gather_cells(self.grad, self.cells64, gb, g.array)
            stats = m.integrate_into(g.array, f.array, None, c.array, dt)
            _scatter_cells(self.flux, self.cells64, fb, f.array)
            _scatter_cells(self.jac, self.cells64, cb, c.array)
            return stats

        # pipelined: while the device works on chunk k (integrate_range_into releases the GIL), the helper thread
        # scatters chunk k-1 and gathers chunk k+1 on the library's host pool
        def gather(k):
            c0, nc = self._chunks[k]
            g = self._stage[k % 2][0].array
            _gather_cells(self.grad, self.cells64[c0:c0 + nc], gb, g[: nc * q])

        def scatter(k):
            c0, nc = self._chunks[k]
            _, f, c = self._stage[k % 2]
            _scatter_cells(self.flux, self.cells64[c0:c0 + nc], fb, f.array[: nc * q])
            _scatter_cells(self.jac, self.cells64[c0:c0 + nc], cb, c.array[: nc * q])

        total = None
        scattered = []
        pending_gather = self._host.submit(gather, 0)
        for k, (c0, nc) in enumerate(self._chunks):
            pending_gather.result()
            if k >= 2:
                scattered[k - 2].result()  # the last reader of this chunk's output buffers
            if k + 1 < len(self._chunks):
                pending_gather = self._host.submit(gather, k + 1)  # its buffer was last read by integrate(k-1): done
            g, f, c = self._stage[k % 2]
            npt = nc * q
            st = m.integrate_range_into(c0 * q, npt, g.array[:npt], f.array[:npt], None, c.array[:npt], dt)
            scattered.append(self._host.submit(scatter, k))
            if total is None:
                total = st
            else:
                total.n_points += st.n_points
                total.n_plastic += st.n_plastic
                total.n_fail += st.n_fail
                total.max_iter = max(total.max_iter, st.max_iter)
                total.max_residual = max(total.max_residual, st.max_residual)
                total.kernel_ms += st.kernel_ms
        for fut in scattered[-2:]:
            fut.result()
        m.last_stats = total
        return total

    def fetch_state(self, flux=True, internal=True):
        """Copy the material's current ``s1`` flux and / or internal state variables into their Functions."""
        m = self.material
        fields = ([(self.fname, self.fdim)] if flux else []) + (list(m.internal_state_variables.items()) if internal else [])
        if self.identity:
            for k, _ in fields:
                m.read_state_into(k, self.flux if k == self.fname else self.isv[k])
        else:
            # only what the Functions hold, through a page-locked staging array
            w = self._stage_state
            for k, d in fields:
                d = max(1, d)
                buf = w.array.reshape(-1)[: self.n * d].reshape(self.n, d)
                m.read_state_into(k, buf)
                _scatter_cells(self.flux if k == self.fname else self.isv[k], self.cells64, self.num_qp * d, buf)

    # ---- quadrature_map.py:350-360 -----------------------------------------------------------------------
    def advance(self):
        self.material.data_manager.update()
        self.fetch_state()
