"""Drop-in ``QuadratureMap`` for CUDA materials: the reference class with its material-facing half routed through
:class:`~dolfinx_materials_b200.exchange.QuadratureExchange`.

    from dolfinx_materials_b200.quadrature_map import gpu_quadrature_map
    QuadratureMap = gpu_quadrature_map()            # subclass of dolfinx_materials.quadrature_map.QuadratureMap
    qmap = QuadratureMap(domain, deg_quad, jm.CUDAMaterial(behavior))      # everything else as in the demos

Only three methods are replaced (reference ``dolfinx_materials/quadrature_map.py``):

* ``update()`` (``:297-334``): external state variables and the gradient expressions are evaluated exactly as before
  (``QuadratureExpression.eval``, ``quadrature_function.py:45-51``); then, instead of gather -> ``integrate`` -> three
  NaN scans -> fancy-index scatters, the exchange hands the gradient Function's ``x.array`` to the library and the
  device writes flux and tangent straight into the ``x.array`` of their Functions (cell subsets: host-pool gather /
  scatter pipelined against the device).  Internal state variables stay on the GPU (unless
  ``internal_state_every_update`` asks for the reference's behaviour).
* ``advance()`` (``:350-360``): ``data_manager.update()`` (a generation swap) and one fetch of flux + internal state
  into their Functions.
* ``initialize_state()`` (``:281-295``): same content, pushed through the exchange.

Everything else -- forms, ``derivative``, ``register_gradient``, ``update_initial_state``, ``project_on``, material
property evaluation -- is inherited untouched.  The same timer names are used (``quadrature_map.py:302-331``).

dolfinx / UFL are not importable in the build container, so ``gpu_quadrature_map(base)`` takes the base class as an
argument; the tests pass a stand-in that exposes the attributes this module touches (``tests/test_quadrature_map_adapter*.py``).
With no argument it imports the real one.
"""

import contextlib

from .exchange import QuadratureExchange

try:
    from dolfinx.common import Timer as _Timer
except Exception:  # noqa: BLE001 - dolfinx is optional here

    def _Timer(name):
        return contextlib.nullcontext()


def gpu_quadrature_map(base=None, strict=True, internal_state_every_update=False):
    """Return a subclass of ``base`` (default: ``dolfinx_materials.quadrature_map.QuadratureMap``) whose ``update`` /
    ``advance`` / ``initialize_state`` go through :class:`QuadratureExchange`.  ``strict``: failed Gauss points raise
    (the reference asserts on NaN, ``quadrature_map.py:322-324``) instead of warning.
    ``internal_state_every_update``: refresh the internal-state Functions on every ``update()`` as the reference does
    (``quadrature_map.py:333``) -- needed only when a form reads them during the Newton iterations; by default they are
    refreshed once per load step, in ``advance()``, and stay on the GPU in between."""
    if base is None:
        from dolfinx_materials.quadrature_map import QuadratureMap as base  # the reference

    class GPUQuadratureMap(base):
        _xchg = None

        def _exchange(self):
            if self._xchg is None:
                mat = self.material
                if mat.rotation_matrix is not None:
                    # rotate_gradients / rotate_fluxes / rotate_tangent_operator (quadrature_map.py:315-330,
                    # mfront.py:336-343) are identities for the isotropic CUDA behaviours: nothing to rotate
                    raise NotImplementedError("CUDA materials are isotropic: rotation_matrix must be None")
                missing = [g for g in mat.gradients if g not in self.gradients]
                if missing:
                    raise ValueError(f"gradient(s) {missing} have not been registered (register_gradient)")
                nqp = len(self.quadrature_points)
                gname, gdim = next(iter(mat.gradients.items()))
                gfun = self.gradients[gname].function
                mesh_cells = gfun.x.array.size // (nqp * gdim)  # the Functions span every cell of the mesh
                self._xchg = QuadratureExchange(
                    mat, mesh_cells, nqp, {gname: gfun}, self.fluxes, self.internal_state_variables,
                    self.jacobian_flatten, cells=self.cells, strict=strict, keep_data_manager=True)
                self._xchg._initialized = self._initialized
            return self._xchg

        def _eval_gradients(self):
            for name in self.material.gradients:
                self.gradients[name].eval(self.cells)

        def initialize_state(self):
            x = self._exchange()
            self._eval_gradients()
            x.initialize_state()
            self._initialized = True

        def update(self):
            x = self._exchange()
            with _Timer("dx_mat: External state variable update"):
                self.update_external_state_variables()
            with _Timer("dx_mat: Gradients evaluation"):
                self._eval_gradients()
            if not self._initialized:
                x.initialize_state()
                self._initialized = True
            x._initialized = True
            with _Timer("dx_mat: Material integration"):
                self.last_stats = x.update(fetch_internal_state=internal_state_every_update)
            return self.last_stats

        def advance(self):
            self._exchange().advance()

        def close(self):
            """Release the page-locked registrations of the Function arrays now (otherwise a finalizer of the exchange
            does it when the map is garbage-collected)."""
            if self._xchg is not None:
                self._xchg.close()
                self._xchg = None

    GPUQuadratureMap.__name__ = "GPUQuadratureMap"
    return GPUQuadratureMap
