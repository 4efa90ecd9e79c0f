"""Stand-ins for the reference classes the drop-in ``GPUQuadratureMap`` builds on, for an environment without dolfinx:

* ``StandInQuadratureMap``: the attributes and methods of ``dolfinx_materials.quadrature_map.QuadratureMap`` that the
  adapter touches or inherits, with the reference's semantics on plain arrays -- ``cells`` / ``dofs`` (``:64-70,
  255-260``), Function-like ``fluxes`` / ``internal_state_variables`` / ``jacobian_flatten`` with ``.x.array``,
  ``gradients[name]`` with ``.function`` and ``.eval(cells)`` (``quadrature_function.py:24-51``),
  ``update_initial_state`` (``:262-279``), ``initialize_state`` (``:281-295``), ``update`` (``:297-334``),
  ``advance`` (``:350-360``).  Run unmodified it IS the reference sequence (the comparison arm of the tests).
* ``StandInMaterial``: a numpy material implementing the protocol methods both arms call (a test double, not a CPU
  fallback of the product).
"""
import numpy as np

from dolfinx_materials_b200.material import IntegrationStats


class _X:
    def __init__(self, n):
        self.array = np.zeros(n)


class Fun:
    """fem.Function on a Quadrature space: flat ``x.array`` of (all mesh points) x dim."""

    def __init__(self, ntot, dim):
        self.dim = max(1, dim)
        self.x = _X(ntot * self.dim)


def _get_vals(fun):
    return fun.x.array.reshape(-1, fun.dim)


def _update_vals(fun, array, cells):
    arr = np.asarray(array).ravel()
    bs = len(arr) // len(cells)
    dofs = np.add.outer(np.asarray(cells) * bs, np.arange(bs)).ravel()
    fun.x.array[dofs] = arr


class Expr:
    """QuadratureExpression: ``eval(cells)`` writes the current value of the expression for those cells."""

    def __init__(self, name, provider, ntot, dim, nqp):
        self.name, self.provider, self.nqp = name, provider, nqp
        self.function = Fun(ntot, dim)

    def eval(self, cells):
        vals = np.asarray(self.provider()).reshape(-1, self.function.dim)
        rows = (self.nqp * np.asarray(cells)[:, None] + np.arange(self.nqp)[None, :]).ravel()
        _get_vals(self.function)[rows] = vals[rows]


class StandInQuadratureMap:
    def __init__(self, num_mesh_cells, nqp, material, cells=None):
        self._nqp = nqp
        self.material = material
        ntot = num_mesh_cells * nqp
        self._ntot = ntot
        if cells is None:
            self.num_cells = num_mesh_cells
            self.cells = np.arange(num_mesh_cells, dtype=np.int32)
        else:
            self.num_cells = len(cells)
            self.cells = np.asarray(cells, dtype=np.int32)
        self.gradients = {}
        buff = sum(int(np.prod(s)) for s in material.tangent_blocks.values())
        self.jacobian_flatten = Fun(ntot, buff)
        self.fluxes = {k: Fun(ntot, d) for k, d in material.fluxes.items()}
        self.internal_state_variables = {k: Fun(ntot, d) for k, d in material.internal_state_variables.items()}
        self.external_state_variables = {}
        self.dofs = self._cell_to_dofs(self.cells)
        self.material.set_data_manager(len(self.dofs))
        for name, prop in self.material.material_properties.items():
            self.material.update_material_property(name, np.asarray(prop))
        self._initialized = False

    @property
    def quadrature_points(self):
        return list(range(self._nqp))

    @property
    def variables(self):
        return {**self.gradients, **self.fluxes, **self.internal_state_variables}

    def _cell_to_dofs(self, cells):
        return (np.repeat(self._nqp * cells[:, None], self._nqp, axis=1) + np.arange(self._nqp)[None, :]).ravel()

    def register_gradient(self, name, provider):
        if name not in self.material.gradients:
            raise ValueError(f"Gradient '{name}' is not available from the material law.")
        self.gradients[name] = Expr(name, provider, self._ntot, self.material.gradients[name], self._nqp)

    def update_external_state_variables(self):
        pass

    def get_gradient_vals(self, gradient, cells):
        gradient.eval(cells)
        return _get_vals(gradient.function)[self.dofs, :]

    def update_initial_state(self, field_name, value=None):
        field = self.fluxes[field_name] if field_name in self.fluxes else self.internal_state_variables[field_name]
        values = _get_vals(field)[self.dofs]
        if value is not None:
            values = np.full_like(values, value)
            _update_vals(field, values, self.cells)
        self.material.set_initial_state_dict({field_name: values})

    def initialize_state(self):
        state = {k: self.get_gradient_vals(f, self.cells) for k, f in self.gradients.items()}
        state.update({k: _get_vals(f)[self.dofs] for k, f in self.fluxes.items()})
        state.update({k: _get_vals(f)[self.dofs] for k, f in self.internal_state_variables.items()})
        self.material.set_initial_state_dict(state)
        self._initialized = True

    def update(self):
        if not self._initialized:
            self.initialize_state()
        self.update_external_state_variables()
        grad_vals = np.concatenate([self.get_gradient_vals(self.gradients[n], self.cells) for n in self.material.gradients], axis=1)
        flux_vals, isv_vals, Ct_vals = self.material.integrate(grad_vals)
        assert not np.any(np.isnan(flux_vals)) and not np.any(np.isnan(isv_vals)) and not np.any(np.isnan(Ct_vals))
        buff = 0
        for name, dim in self.material.fluxes.items():
            _update_vals(self.fluxes[name], flux_vals[:, buff: buff + dim], self.cells)
            buff += dim
        buff = 0
        for name, dim in self.material.internal_state_variables.items():
            _update_vals(self.internal_state_variables[name], isv_vals[:, buff: buff + dim], self.cells)
            buff += dim
        _update_vals(self.jacobian_flatten, Ct_vals, self.cells)

    def advance(self):
        self.material.data_manager.update()
        final_state = self.material.get_final_state_dict()
        for key in self.variables:
            if key not in self.gradients:
                _update_vals(self.variables[key], final_state[key], self.cells)


class _DataManager:
    def __init__(self, m):
        self.m = m

    def update(self):
        self.m.s0 = {k: v.copy() for k, v in self.m.s1.items()}


class StandInMaterial:
    """flux = 2 * strain + p_old, Ct row = outer(strain, 1..6) flattened, p_new = p_old + |strain|_1, epsp = -strain."""

    gradients = {"strain": 6}
    fluxes = {"stress": 6}
    internal_state_variables = {"p": 1, "epsp": 6}
    tangent_blocks = {("stress", "strain"): (6, 6)}
    material_properties = {"E": 1.0}
    rotation_matrix = None

    def __init__(self):
        self.calls = []
        self._n = None

    def set_data_manager(self, n):
        self.n = self._n = n
        self.s0 = {"strain": np.zeros((n, 6)), "stress": np.zeros((n, 6)), "p": np.zeros((n, 1)), "epsp": np.zeros((n, 6))}
        self.s1 = {k: v.copy() for k, v in self.s0.items()}
        self.data_manager = _DataManager(self)

    def update_material_property(self, name, value):
        pass

    def set_initial_state_dict(self, state):
        for k, v in state.items():
            self.s0[k] = np.array(v, dtype=float).reshape(self.n, -1)

    def get_final_state_dict(self):
        return {k: v.copy() for k, v in self.s1.items()}

    def _compute(self, sl, g, flux, ct):
        g = np.asarray(g).reshape(-1, 6)
        p_old = self.s0["p"][sl]
        self.s1["strain"][sl] = g
        self.s1["stress"][sl] = 2.0 * g + p_old
        self.s1["p"][sl] = p_old + np.abs(g).sum(axis=1, keepdims=True)
        self.s1["epsp"][sl] = -g
        if flux is not None:
            flux.reshape(-1, 6)[:] = self.s1["stress"][sl]
        if ct is not None:
            ct.reshape(-1, 36)[:] = (g[:, :, None] * np.arange(1.0, 7.0)[None, None, :]).reshape(-1, 36)
        return IntegrationStats(n_points=len(g), n_plastic=int((g[:, 0] > 0).sum()), max_iter=3, max_residual=float(np.abs(g).max()))

    def integrate(self, g, dt=0):
        n = self.n
        flux, ct = np.empty((n, 6)), np.empty((n, 36))
        self.calls.append(("integrate", 0, n))
        self._compute(slice(0, n), g, flux, ct)
        return flux, np.concatenate([self.s1["p"], self.s1["epsp"]], axis=1), ct.reshape(n, 6, 6)

    def integrate_into(self, g, flux_out=None, isv_out=None, ct_out=None, dt=0):
        self.calls.append(("all", 0, self.n))
        return self._compute(slice(0, self.n), g, flux_out, ct_out)

    def integrate_range_into(self, start, count, g, flux_out=None, isv_out=None, ct_out=None, dt=0):
        assert start % 2 == 0
        self.calls.append(("range", start, count))
        return self._compute(slice(start, start + count), g, flux_out, ct_out)

    def read_state_into(self, key, out, gen=1):
        out.reshape(self.n, -1)[:] = (self.s1 if gen == 1 else self.s0)[key]
