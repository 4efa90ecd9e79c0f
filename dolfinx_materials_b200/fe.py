"""GPU gradient evaluation: the device-side counterpart of
``QuadratureExpression.eval`` + ``QuadratureMap.get_gradient_vals`` (reference
``dolfinx_materials/quadrature_function.py:45-51``, ``quadrature_map.py:251-253``) for displacement-gradient
expressions on affine simplex meshes.  Only the displacement vector is sent to the device; the gradients are
written straight into the material's SoA gradient buffer, ready for ``integrate_resident``.

With dolfinx the constructor arguments are ``mesh.geometry.x``, ``mesh.geometry.dofmap``, ``V.dofmap.list``
and ``basix`` tabulated first derivatives at the quadrature points (``element.tabulate(1, points)[1:]``
transposed to ``(nqp, ndofs, tdim)``); nothing here imports dolfinx.
"""

import ctypes
import weakref

import numpy as np

from . import _lib
from ._lib import MEM_HOST, check

KIND_STRAIN, KIND_DEFGRAD = 0, 1


class GradientEvaluator:
    def __init__(self, material, coords, geom_dofmap, u_dofmap, dphi, tdim=3):
        lib = _lib.load()
        material._require_handle()
        self.material = material
        self.tdim = int(tdim)
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        if coords.ndim != 2 or coords.shape[1] != 3:
            raise ValueError("coords must be (num_nodes, 3) like mesh.geometry.x")
        gd = np.ascontiguousarray(geom_dofmap, dtype=np.int32)
        ud = np.ascontiguousarray(u_dofmap, dtype=np.int32)
        dphi = np.ascontiguousarray(dphi, dtype=np.float64)
        if gd.shape[1] != self.tdim + 1:
            raise ValueError("affine simplex cells only: geom_dofmap must be (num_cells, tdim+1)")
        if dphi.shape[1:] != (ud.shape[1], self.tdim) or gd.shape[0] != ud.shape[0]:
            raise ValueError("dphi must be (nqp, ndofs_cell, tdim) and the dofmaps must cover the same cells")
        self.num_cells, self.nqp = ud.shape[0], dphi.shape[0]
        self.num_dofs = int(ud.max()) + 1
        self.kind = KIND_DEFGRAD if material.behavior.finite_strain else KIND_STRAIN
        h = ctypes.c_void_p()
        check(
            lib.dxm_mesh_create(material.device, self.tdim, self.num_cells, coords.shape[0], coords.ctypes.data_as(ctypes.c_void_p),
                                gd.ctypes.data_as(ctypes.c_void_p), ud.shape[1], ud.ctypes.data_as(ctypes.c_void_p),
                                self.num_dofs, self.nqp, dphi.ctypes.data_as(ctypes.c_void_p), ctypes.byref(h)),
            "dxm_mesh_create",
        )
        self._h = h
        self._fin = weakref.finalize(self, lib.dxm_mesh_destroy, h)

    def eval(self, u):
        """``u``: blocked displacement vector (``u.x.array``), length ``num_dofs * tdim``."""
        u = np.ascontiguousarray(u, dtype=np.float64).ravel()
        if u.size != self.num_dofs * self.tdim:
            raise ValueError(f"u must have {self.num_dofs * self.tdim} entries, got {u.size}")
        check(_lib.load().dxm_eval_gradient(self._h, self.material._h, u.ctypes.data_as(ctypes.c_void_p), MEM_HOST, self.kind),
              "dxm_eval_gradient")
