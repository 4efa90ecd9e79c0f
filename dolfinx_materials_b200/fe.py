"""GPU gradient evaluation: the device-side counterpart of
``QuadratureExpression.eval`` + ``QuadratureMap.get_gradient_vals`` (reference
``dolfinx_materials/quadrature_function.py:45-51``, ``quadrature_map.py:251-253``) for displacement-gradient
expressions on affine simplex meshes.  Only the displacement vector is sent to the device; the gradients are
written straight into the material's SoA gradient buffer, ready for ``integrate_resident``.

``ElementForms`` / ``AssembledSystem`` are the consumers on the other side of the update:
the element residual vectors / tangent matrices that DOLFINx's cell kernels compute from the flux and
``jacobian_flatten`` Quadrature Functions (``solvers.py:80-81``, ``quadrature_map.py:132-158``), formed on the
device from the SoA outputs of the last ``integrate`` -- either handed out per element (``MatSetValuesLocal``
input) or scatter-added into a device-resident CSR system so that the tangent never crosses PCIe.

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
    def __init__(self, material, coords, geom_dofmap, u_dofmap, dphi, tdim=3, num_dofs=None):
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
        self.num_cells, self.nqp, self.ndofs_cell = ud.shape[0], dphi.shape[0], ud.shape[1]
        # num_dofs: size of the (global) space when this evaluator holds only a block of its cells (one rank's share)
        self.num_dofs = int(ud.max()) + 1 if num_dofs is None else int(num_dofs)
        if self.num_dofs <= int(ud.max()):
            raise ValueError("num_dofs is smaller than the largest dof index of the dofmap")
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


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class ElementForms:
    """Element vectors ``fe (num_cells, nd*tdim)`` and matrices ``ke (num_cells, nd*tdim, nd*tdim)`` of
    ``dot(flux, dgrad(v))*dx`` and ``dgrad(v).Ct.dgrad(du)*dx`` from the material's last update.
    ``weights``: reference-cell quadrature weights (``basix.make_quadrature(celltype, degree)[1]``)."""

    def __init__(self, evaluator, weights):
        self.ev = evaluator
        w = np.ascontiguousarray(weights, dtype=np.float64).ravel()
        if w.size != evaluator.nqp:
            raise ValueError(f"expected {evaluator.nqp} quadrature weights, got {w.size}")
        check(_lib.load().dxm_mesh_set_weights(evaluator._h, _ptr(w)), "dxm_mesh_set_weights")
        self.ndof = evaluator.ndofs_cell * evaluator.tdim

    def compute(self, vector=True, matrix=True, out_fe=None, out_ke=None):
        ev = self.ev
        fe = ke = None
        if vector:
            fe = out_fe if out_fe is not None else np.empty((ev.num_cells, self.ndof))
        if matrix:
            ke = out_ke if out_ke is not None else np.empty((ev.num_cells, self.ndof, self.ndof))
        for arr in (fe, ke):
            if arr is not None and (arr.dtype != np.float64 or not arr.flags.c_contiguous):
                raise TypeError("outputs must be C-contiguous float64 arrays")
        check(_lib.load().dxm_element_forms(ev._h, ev.material._h, ev.kind, _ptr(fe), _ptr(ke), MEM_HOST),
              "dxm_element_forms")
        return fe, ke


class AssembledSystem:
    """Device-resident CSR tangent + residual vector of the blocked space (pattern from DOLFINx
    ``create_matrix`` / PETSc ``getRowIJ``); ``bc``: optional boolean marker of constrained global dofs."""

    def __init__(self, forms, rowptr, colidx, bc=None):
        lib = _lib.load()
        self.forms = forms
        ev = forms.ev
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        self.colidx = np.ascontiguousarray(colidx, dtype=np.int32)
        self.nrows = self.rowptr.size - 1
        if self.nrows != ev.num_dofs * ev.tdim:
            raise ValueError(f"pattern has {self.nrows} rows, the space has {ev.num_dofs * ev.tdim} dofs")
        h = ctypes.c_void_p()
        check(lib.dxm_system_create(ev.material.device, self.nrows, _ptr(self.rowptr), _ptr(self.colidx), ctypes.byref(h)),
              "dxm_system_create")
        self._h = h
        self._fin = weakref.finalize(self, lib.dxm_system_destroy, h)
        self.nnz = int(lib.dxm_system_nnz(h))
        self._pin_vals = self._pin_rhs = None
        if bc is not None:
            self.set_bc(bc)

    def set_bc(self, marker):
        m = None if marker is None else np.ascontiguousarray(np.asarray(marker, dtype=bool).astype(np.uint8))
        if m is not None and m.size != self.nrows:
            raise ValueError("bc marker must have one entry per global dof")
        check(_lib.load().dxm_system_set_bc(self._h, _ptr(m)), "dxm_system_set_bc")

    def set_lifting(self, values):
        """Prescribed solution values on the constrained dofs (``None``: homogeneous): ``assemble`` then forms
        ``rhs -= A[:, bc] x_bc`` and ``rhs[bc] = x_bc`` (``apply_lifting`` + ``set_bc``)."""
        v = None if values is None else np.ascontiguousarray(values, dtype=np.float64).ravel()
        if v is not None and v.size != self.nrows:
            raise ValueError("lifting values must have one entry per global dof")
        check(_lib.load().dxm_system_set_lifting(self._h, _ptr(v)), "dxm_system_set_lifting")

    def assemble(self, vector=True, matrix=True):
        ev = self.forms.ev
        check(_lib.load().dxm_assemble(ev._h, ev.material._h, ev.kind, self._h, int(vector), int(matrix)), "dxm_assemble")

    # ---- rank-sharded assembly (one process per GPU, each with a contiguous block of cells) -----------------
    def device_arrays(self):
        """Zero-copy ``torch`` views of the CSR value array (nnz,) and the right-hand side (nrows,) on the device."""
        import torch

        pv, pr = ctypes.c_void_p(), ctypes.c_void_p()
        check(_lib.load().dxm_system_device_ptrs(self._h, ctypes.byref(pv), ctypes.byref(pr)), "dxm_system_device_ptrs")

        class _Cai:  # __cuda_array_interface__ carrier
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}

        dev = torch.device("cuda", self.forms.ev.material.device)
        return (torch.as_tensor(_Cai(pv.value, self.nnz), device=dev), torch.as_tensor(_Cai(pr.value, self.nrows), device=dev))

    def assemble_sharded(self, group=None, vector=True, matrix=True):
        """Assemble this rank's cells, sum values / rhs over the process group (NCCL all-reduce over NVLink -- the
        one real exchange step of the FE side, what PETSc's assembly does for the reference), then apply the
        constraints once.  Every rank ends up with the full system."""
        import torch.distributed as dist

        lib = _lib.load()
        check(lib.dxm_system_defer_constraints(self._h, 1), "dxm_system_defer_constraints")
        try:
            self.assemble(vector=vector, matrix=matrix)
        finally:
            check(lib.dxm_system_defer_constraints(self._h, 0), "dxm_system_defer_constraints")
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            vals, rhs = self.device_arrays()
            if matrix:
                dist.all_reduce(vals, group=group)
            if vector:
                dist.all_reduce(rhs, group=group)
            import torch

            torch.cuda.synchronize()
        check(lib.dxm_system_apply_constraints(self._h), "dxm_system_apply_constraints")

    def get(self, values=True, rhs=True):
        """-> (CSR values (nnz,), rhs (nrows,)) on the host (``A.setValuesCSR(rowptr, colidx, values)``).  The arrays
        are page-locked buffers owned by this object and overwritten by the next ``get`` (copy to keep)."""
        from .material import PinnedArray

        if values and self._pin_vals is None:
            self._pin_vals = PinnedArray((self.nnz,))
        if rhs and self._pin_rhs is None:
            self._pin_rhs = PinnedArray((self.nrows,))
        v = self._pin_vals.array if values else None
        b = self._pin_rhs.array if rhs else None
        check(_lib.load().dxm_system_get(self._h, _ptr(v), _ptr(b), MEM_HOST), "dxm_system_get")
        return v, b

    def solve(self, rtol=1e-8, maxit=20000):
        """``A x = rhs`` on the device (BiCGStab + block-Jacobi); returns ``(x, iterations, relative residual,
        converged)``.  Stand-in for the reference's PETSc KSP in DOLFINx-free runs, not part of the drop-in path."""
        x = np.empty(self.nrows)
        it, rel = ctypes.c_int(0), ctypes.c_double(0.0)
        rc = check(_lib.load().dxm_system_solve(self._h, self.forms.ev.tdim, float(rtol), int(maxit), _ptr(x), MEM_HOST,
                                                ctypes.byref(it), ctypes.byref(rel)), "dxm_system_solve")
        return x, it.value, rel.value, rc == 0
