"""Protocol-replay harness: the call sequence and array layouts of the reference's ``QuadratureMap``
(``dolfinx_materials/quadrature_map.py``) on plain numpy arrays standing in for dolfinx Quadrature
``Function.x.array`` vectors.  dolfinx/UFL are not installable here, so this is how the drop-in
boundary is exercised exactly as the reference's caller exercises it:

* ``__init__``                    quadrature_map.py:51-130  (cells, dofs = num_qp*cell + q :255-260,
                                   set_data_manager(len(dofs)) :231-233, update_material_properties :160-172)
* ``initialize_state``            :281-295  (gradients + fluxes + isv -> set_initial_state_dict)
* ``update_initial_state``        :262-279
* ``update``                      :297-334  (gather gradients [dofs,:], concatenate, integrate, NaN asserts,
                                   scatter flux / isv column blocks / flattened tangent with _update_vals)
* ``advance``                     :350-360  (data_manager.update(), then get_final_state_dict() scatter)
* ``_get_vals`` / ``_update_vals`` utils.py:98-104, :136-143
"""
import numpy as np


class _Function:
    def __init__(self, ntot, dim):
        self.dim = max(1, dim)
        self.array = np.zeros(ntot * self.dim)


def _get_vals(fun):
    return fun.array.reshape((-1, fun.dim))


def _update_vals(fun, array, cells=None):
    if cells is None:
        fun.array[:] = array.ravel()
    else:
        arr = np.asarray(array).ravel()
        bs = len(arr) // len(cells)
        dofs = np.add.outer(cells * bs, np.arange(bs)).ravel()
        fun.array[dofs] = arr


class QuadratureMapReplay:
    def __init__(self, num_cells, num_qp, material, cells=None):
        self.num_cells, self.num_qp, self.material = num_cells, num_qp, material
        self.cells = np.arange(num_cells, dtype=np.int32) if cells is None else np.asarray(cells, dtype=np.int32)
        ntot = num_cells * num_qp
        buff = sum(nf * ng for (nf, ng) in material.tangent_blocks.values())
        self.jacobian_flatten = _Function(ntot, buff)
        self.fluxes = {k: _Function(ntot, d) for k, d in material.fluxes.items()}
        self.internal_state_variables = {k: _Function(ntot, d) for k, d in material.internal_state_variables.items()}
        self.gradients = {}  # name -> _Function holding the "evaluated UFL expression"
        self._initialized = False
        self.dofs = (np.repeat(num_qp * self.cells[:, None], num_qp, axis=1) + np.arange(num_qp)[None, :]).ravel()
        self.material.set_data_manager(len(self.dofs))
        assert material.rotation_matrix is None
        for name, prop in material.material_properties.items():
            values = np.asarray(prop)
            self.material.update_material_property(name, values)

    @property
    def variables(self):
        return {**self.gradients, **self.fluxes, **self.internal_state_variables}

    def register_gradient(self, name, values_all_points):
        """``values_all_points``: (num_cells*num_qp, dim) -- what fem.Expression.eval would scatter."""
        if name not in self.material.gradients:
            raise ValueError(f"Gradient '{name}' is not available from the material law.")
        f = _Function(self.num_cells * self.num_qp, self.material.gradients[name])
        f.array[:] = np.asarray(values_all_points).ravel()
        self.gradients[name] = f

    def set_gradient_values(self, name, values_all_points):
        self.gradients[name].array[:] = np.asarray(values_all_points).ravel()

    def update_initial_state(self, field_name, value):
        field = self.variables[field_name]
        values = _get_vals(field)[self.dofs]
        values = np.full_like(values, value)
        _update_vals(field, values, self.cells)
        self.material.set_initial_state_dict({field_name: values})

    def initialize_state(self):
        state_flux = {k: _get_vals(f)[self.dofs] for k, f in self.fluxes.items()}
        state_isv = {k: _get_vals(f)[self.dofs] for k, f in self.internal_state_variables.items()}
        state_grad = {k: _get_vals(f)[self.dofs, :] for k, f in self.gradients.items()}
        self.material.set_initial_state_dict({**state_grad, **state_flux, **state_isv})
        self._initialized = True

    def update(self):
        if not self._initialized:
            self.initialize_state()
        grad_vals = [_get_vals(self.gradients[name])[self.dofs, :] for name in self.material.gradients.keys()]
        grad_vals = np.concatenate(grad_vals, axis=1)
        flux_vals, isv_vals, Ct_vals = self.material.integrate(grad_vals)
        assert not (np.any(np.isnan(flux_vals)))
        assert not (np.any(np.isnan(isv_vals)))
        assert not (np.any(np.isnan(Ct_vals)))
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
        for key in self.variables.keys():
            if key not in self.gradients:
                _update_vals(self.variables[key], final_state[key], self.cells)
