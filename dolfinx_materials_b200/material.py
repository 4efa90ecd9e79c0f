"""``CUDAMaterial``: the CUDA-backed ``Material`` of the reference protocol.

Mirrors, name for name, what ``QuadratureMap`` and the solvers call on a material
(reference ``dolfinx_materials/generic.py:103-201`` and ``jaxmat.py:141-234``; call sites
``quadrature_map.py:84-137,160-172,231-233,279-360``): ``gradients`` / ``fluxes`` /
``internal_state_variables`` / ``tangent_blocks`` / ``variables`` / ``*_names`` / ``rotation_matrix`` /
``material_properties`` / ``update_material_property`` / ``set_data_manager`` / ``integrate`` /
``get_initial_state_dict`` / ``get_final_state_dict`` / ``set_initial_state_dict`` /
``data_manager.update()`` / ``data_manager.revert()``.

State lives on the device in SoA buffers owned by ``libdxm_cuda.so``; ``integrate`` hands back host
``(flux, isv, Ct)`` arrays exactly like the reference, and the ``*_resident`` methods expose the
zero-copy device path (DLPack) for callers that keep gradients on the GPU.
"""

import ctypes
import warnings
import weakref
from dataclasses import dataclass

import numpy as np

from . import PerformanceWarning, _lib
from ._lib import MEM_HOST, MEM_RESIDENT, Stats, check


try:  # same timer names as the reference when dolfinx is there (jaxmat.py:209-223, quadrature_map.py:320)
    from dolfinx.common import Timer as _Timer
except Exception:  # noqa: BLE001 - dolfinx is optional
    import contextlib

    def _Timer(name):
        return contextlib.nullcontext()


def _sym6_packed(c):
    j, i = sorted(divmod(c, 6))
    return j * 6 - (j * (j - 1)) // 2 + (i - j)


# row of the packed resident tangent that holds entry c = j*6+i of the row-major symmetric 6x6 (include/dxm.h)
SYM6_PACKED = np.array([_sym6_packed(c) for c in range(36)])


@dataclass
class IntegrationStats:
    """Per-call statistics reduced on the device (fused replacement of the host NaN scans,
    reference ``quadrature_map.py:322-324``)."""

    n_points: int = 0
    n_plastic: int = 0
    n_fail: int = 0
    max_iter: int = 0
    max_residual: float = 0.0
    kernel_ms: float = 0.0

    @classmethod
    def from_c(cls, s):
        return cls(s.n_points, s.n_plastic, s.n_fail, s.max_iter, s.max_residual, s.kernel_ms)


class _PinnedBlock:
    """Owner of one page-locked allocation (``dxm_host_alloc``): freed when the last reference goes -- the
    :class:`_Pinned` wrapper or ANY numpy view handed out from it (the views keep the block alive through their base)."""

    def __init__(self, nbytes):
        lib = _lib.load()
        ptr = ctypes.c_void_p()
        check(lib.dxm_host_alloc(ctypes.byref(ptr), int(nbytes)), "dxm_host_alloc")
        self.ptr = ptr.value
        self._fin = weakref.finalize(self, lib.dxm_host_free, ctypes.c_void_p(self.ptr))


class _Pinned:
    """A page-locked host array viewed as numpy.  The array (and every view sliced from it, e.g. what ``integrate``
    returns) holds a reference to the allocation, so dropping the material or re-creating its data manager never
    leaves a caller with a dangling view (the reference returns ordinary owned arrays, ``generic.py:185-189``)."""

    def __init__(self, shape):
        size = int(np.prod(shape))
        self.block = _PinnedBlock(size * 8)
        self.ptr = self.block.ptr
        buf = (ctypes.c_double * max(size, 1)).from_address(self.ptr)
        buf._dxm_owner = self.block  # numpy keeps `buf` as the base of the array: the block lives as long as any view
        self.array = np.ctypeslib.as_array(buf)[:size].reshape(shape)


def pin_array(arr):
    """Page-lock a caller-owned ndarray in place (``dxm_host_register``); returns a callable that
    releases it.  Long-lived arrays only (registration costs ~0.2 ms / MB)."""
    lib = _lib.load()
    if not arr.flags.c_contiguous:
        raise ValueError("only C-contiguous arrays can be page-locked")
    addr = ctypes.c_void_p(arr.ctypes.data)
    check(lib.dxm_host_register(addr, arr.nbytes), "dxm_host_register")
    return lambda: lib.dxm_host_unregister(addr)


PinnedArray = _Pinned  # public name: ``PinnedArray(shape).array`` is a page-locked ndarray view


class _StateView:
    """Dict-like view of one state generation (``s0`` / ``s1``); values are fetched from the device
    on access, shape ``(n, dim)`` like ``MaterialStateManager.__getitem__`` (``generic.py:260-271``)."""

    def __init__(self, material, gen):
        self._m = material
        self._gen = gen

    def keys(self):
        return list(self._m.variables.keys())

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self.keys())

    def __contains__(self, key):
        return key in self._m.variables

    def __getitem__(self, key):
        if isinstance(key, slice):  # generic.py:195 uses s0[:]
            return {k: self._m._get_state(self._gen, k) for k in self.keys()}
        return self._m._get_state(self._gen, key)

    def get(self, key, default=None):
        return self[key] if key in self else default

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def values(self):
        return [self[k] for k in self.keys()]


class DeviceDataManager:
    """Device-resident counterpart of ``DataManager`` (``generic.py:204-216``, ``jaxmat.py:30-43``):
    two state generations, ``update()`` = s0 <- s1 and ``revert()`` = s1 <- s0, both O(1) swaps."""

    def __init__(self, material, ngauss):
        self._m = material
        self.n = int(ngauss)
        num_gradients = sum(material.gradients.values())
        num_fluxes = sum(material.fluxes.values())
        self.K = np.zeros((num_fluxes, num_gradients))
        self.s0 = _StateView(material, 0)
        self.s1 = _StateView(material, 1)

    def update(self):
        check(_lib.load().dxm_update(self._m._h), "dxm_update")

    def revert(self):
        check(_lib.load().dxm_revert(self._m._h), "dxm_revert")


class CUDAMaterial:
    """Converts a behaviour descriptor into a dolfinx_materials-compatible material running on a
    B200 (the role ``JAXMaterial(behavior)`` plays in the reference, ``jaxmat.py:141-156``)."""

    def __init__(self, behavior, jit=True, device=None, warn_on_failure=True):
        """``jit``: accepted for signature compatibility with ``JAXMaterial(behavior, jit=True)`` (``jaxmat.py:144``) and
        ignored -- the kernels are compiled ahead of time.  ``device``: CUDA device index; ``None`` = this process's
        local rank modulo the number of devices (one MPI rank / torchrun worker per GPU), 0 for a single process."""
        del jit
        self.behavior = behavior
        if device is None:
            from .distributed import default_device

            device = default_device()
        self.device = int(device)
        self.warn_on_failure = warn_on_failure
        self.material_properties = dict(behavior.properties())
        self._h = None
        self._n = 0
        self._out = None
        self._fin = None
        self.data_manager = None
        self.last_stats = IntegrationStats()

    # ---- protocol: description ---------------------------------------------------------------
    @property
    def name(self):
        return self.behavior.__class__.__name__

    @property
    def rotation_matrix(self):
        return None

    # The reference rotates gradients into the material frame and fluxes / tangent back out of it only when
    # ``rotation_matrix`` is not None (``quadrature_map.py:315-330``; MFront orthotropic behaviours, ``mfront.py:336-343``).
    # Every CUDA behaviour is isotropic -- its response is the same in any frame -- so ``rotation_matrix`` is None, the
    # callers never reach these, and they are identities kept for the completeness of the protocol.
    def rotate_gradients(self, gradient_vals, rotation_values):
        return gradient_vals

    def rotate_fluxes(self, flux_vals, rotation_values):
        return flux_vals

    def rotate_tangent_operator(self, Ct_vals, rotation_values):
        return Ct_vals

    @property
    def gradients(self):
        return {"F": 9} if self.behavior.finite_strain else {"strain": 6}

    @property
    def fluxes(self):
        return {"PK1": 9} if self.behavior.finite_strain else {"stress": 6}

    @property
    def internal_state_variables(self):
        return {"p": 1, "be_bar": 6} if self.behavior.finite_strain else {"p": 1, "epsp": 6}

    @property
    def tangent_blocks(self):
        return {
            (kf, kg): (vf, vg)
            for (kf, vf), (kg, vg) in zip(self.fluxes.items(), self.gradients.items())
        }

    @property
    def variables(self):
        return {**self.gradients, **self.fluxes, **self.internal_state_variables}

    @property
    def gradient_names(self):
        return list(self.gradients.keys())

    @property
    def flux_names(self):
        return list(self.fluxes.keys())

    @property
    def internal_state_variable_names(self):
        return list(self.internal_state_variables.keys())

    # ---- protocol: properties ----------------------------------------------------------------
    def update_material_property(self, key, value):
        """Scalar (0-d) or per-Gauss-point array, as ``QuadratureMap.update_material_properties``
        passes them (``quadrature_map.py:160-172``)."""
        if key not in self.material_properties:
            raise KeyError(f"'{key}' is not a property of {self.name}: {list(self.material_properties)}")
        if self._h is not None:
            self._push_property(key, value)  # raises on a rejected value: the stored property stays as it was
        self.material_properties[key] = value

    def _push_property(self, key, value):
        lib = _lib.load()
        v = np.ascontiguousarray(np.asarray(value, dtype=np.float64).ravel())
        if v.size not in (1, self._n):
            raise ValueError(f"property '{key}' has {v.size} values, expected 1 or {self._n}")
        check(
            lib.dxm_set_property(self._h, key.encode(), v.ctypes.data_as(ctypes.c_void_p), v.size, MEM_HOST),
            "dxm_set_property",
        )

    # ---- protocol: data manager --------------------------------------------------------------
    def set_data_manager(self, ngauss):
        lib = _lib.load()
        if self._fin is not None:
            self._fin()
        h = ctypes.c_void_p()
        check(lib.dxm_create(self.behavior.kind, self.device, int(ngauss), ctypes.byref(h)), "dxm_create")
        self._h = h
        self._n = int(ngauss)
        self._fin = weakref.finalize(self, lib.dxm_destroy, h)
        # the resident call is latency-critical for small batches: bound function and reusable statistics struct
        self._c_integrate = lib.dxm_integrate
        self._c_stats = Stats()
        self._c_stats_ref = ctypes.byref(self._c_stats)
        self._out = None
        self.data_manager = DeviceDataManager(self, ngauss)
        for key, value in self.material_properties.items():
            self._push_property(key, value)
        table = getattr(self.behavior, "yield_stress", None)
        if hasattr(table, "p") and hasattr(table, "sig"):  # TabulatedHardening
            tp = np.ascontiguousarray(table.p, dtype=np.float64).ravel()
            ts = np.ascontiguousarray(table.sig, dtype=np.float64).ravel()
            if tp.size != ts.size:
                raise ValueError("TabulatedHardening: p and sig must have the same length")
            check(lib.dxm_set_hardening_table(self._h, tp.ctypes.data_as(ctypes.c_void_p), ts.ctypes.data_as(ctypes.c_void_p),
                                              int(tp.size)), "dxm_set_hardening_table")

    def _require_handle(self):
        if self._h is None:
            raise RuntimeError("set_data_manager(ngauss) must be called before using the material")

    # ---- protocol: state ---------------------------------------------------------------------
    def _get_state(self, gen, key):
        self._require_handle()
        lib = _lib.load()
        dim = lib.dxm_field_dim(self._h, key.encode())
        if dim < 0:
            raise KeyError(key)
        out = np.empty((self._n, dim))
        check(lib.dxm_get_state(self._h, gen, key.encode(), out.ctypes.data_as(ctypes.c_void_p), MEM_HOST), "dxm_get_state")
        return out

    def _set_state(self, gen, state):
        self._require_handle()
        lib = _lib.load()
        unknown = [k for k in state if k not in self.variables]
        assert len(unknown) == 0, "Material state contains unknown field to update with."
        for key, value in state.items():
            dim = self.variables[key]
            v = np.ascontiguousarray(np.asarray(value, dtype=np.float64).reshape(self._n, dim))
            check(lib.dxm_set_state(self._h, gen, key.encode(), v.ctypes.data_as(ctypes.c_void_p), MEM_HOST), "dxm_set_state")

    def get_initial_state_dict(self):
        return self.data_manager.s0[:]

    def get_final_state_dict(self):
        return self.data_manager.s1[:]

    def set_initial_state_dict(self, state):
        return self._set_state(0, state)

    # ---- protocol: the hot call --------------------------------------------------------------
    def _outputs(self):
        if self._out is None:
            nf = sum(self.fluxes.values())
            ng = sum(self.gradients.values())
            ni = sum(self.internal_state_variables.values())
            self._out = (_Pinned((self._n, nf)), _Pinned((self._n, ni)), _Pinned((self._n, nf, ng)))
        return self._out

    def _finish(self, rc, stats):
        if rc < 0:
            check(rc, "dxm_integrate")
        self.last_stats = IntegrationStats(stats.n_points, stats.n_plastic, stats.n_fail, stats.max_iter,
                                           stats.max_residual, stats.kernel_ms)
        if rc > 0 and self.warn_on_failure:
            warnings.warn(
                f"{rc} Gauss point(s) failed their local constitutive solve "
                f"(max residual {stats.max_residual:.3e})",
                PerformanceWarning,
            )

    def integrate(self, gradients, dt=0):
        """``(flux, isv, Ct) = integrate(gradients, dt)`` with host arrays, reference semantics
        (``generic.py:176-189``, ``jaxmat.py:208-234``): reads s0, writes s1.  The returned arrays are
        page-locked buffers owned by the material and overwritten by the next call (as the reference's
        ``s1.fluxes`` / ``s1.internal_state_variables`` are)."""
        self._require_handle()
        lib = _lib.load()
        ng = sum(self.gradients.values())
        g = np.ascontiguousarray(np.asarray(gradients, dtype=np.float64))
        if g.shape != (self._n, ng):
            raise ValueError(f"gradients must have shape {(self._n, ng)}, got {g.shape}")
        flux, isv, ct = self._outputs()
        stats = Stats()
        with _Timer("dxm: Constitutive update"):
            rc = lib.dxm_integrate(
                self._h, g.ctypes.data_as(ctypes.c_void_p), MEM_HOST, float(dt),
                ctypes.c_void_p(flux.ptr), ctypes.c_void_p(isv.ptr), ctypes.c_void_p(ct.ptr), MEM_HOST,
                ctypes.byref(stats),
            )
        self._finish(rc, stats)
        return flux.array, isv.array, ct.array

    def integrate_into(self, gradients, flux_out=None, isv_out=None, ct_out=None, dt=0):
        """Same update as :meth:`integrate`, but results are written into caller-owned C-contiguous
        float64 arrays (any of them may be ``None`` = not transferred).  With arrays page-locked by
        :func:`pin_array` (e.g. the ``x.array`` of the dolfinx Quadrature Functions) the device DMAs
        straight into them: no staging copy, no scatter.  Returns :class:`IntegrationStats`."""
        self._require_handle()
        lib = _lib.load()
        ng = sum(self.gradients.values())
        nf = sum(self.fluxes.values())
        ni = sum(self.internal_state_variables.values())
        g = np.asarray(gradients)
        if g.dtype != np.float64 or not g.flags.c_contiguous or g.size != self._n * ng:
            raise ValueError(f"gradients must be C-contiguous float64 with {self._n * ng} entries")

        def ptr(a, size, name):
            if a is None:
                return None
            if a.dtype != np.float64 or not a.flags.c_contiguous or a.size != size or not a.flags.writeable:
                raise ValueError(f"{name} must be a writeable C-contiguous float64 array with {size} entries")
            return a.ctypes.data_as(ctypes.c_void_p)

        stats = Stats()
        rc = lib.dxm_integrate(
            self._h, g.ctypes.data_as(ctypes.c_void_p), MEM_HOST, float(dt),
            ptr(flux_out, self._n * nf, "flux_out"), ptr(isv_out, self._n * ni, "isv_out"),
            ptr(ct_out, self._n * nf * ng, "ct_out"), MEM_HOST, ctypes.byref(stats),
        )
        self._finish(rc, stats)
        return self.last_stats

    def integrate_range_into(self, start, count, gradients, flux_out=None, isv_out=None, ct_out=None, dt=0):
        """:meth:`integrate_into` for the points ``[start, start + count)`` only (``start`` even); the arrays hold
        ``count`` rows.  Lets a caller overlap its own host work (gather / scatter of a cell-subset map) with the
        device; every range of a step must be integrated before ``data_manager.update()`` or reading ``s1``.
        Returns the :class:`IntegrationStats` of the range (not stored in ``last_stats``)."""
        self._require_handle()
        lib = _lib.load()
        ng = sum(self.gradients.values())
        nf = sum(self.fluxes.values())
        ni = sum(self.internal_state_variables.values())
        count = int(count)

        def ptr(a, size, name, writeable=True):
            if a is None:
                return None
            if a.dtype != np.float64 or not a.flags.c_contiguous or a.size != size or (writeable and not a.flags.writeable):
                raise ValueError(f"{name} must be a C-contiguous float64 array with {size} entries")
            return a.ctypes.data_as(ctypes.c_void_p)

        stats = Stats()
        rc = lib.dxm_integrate_range(
            self._h, int(start), count, ptr(np.asarray(gradients), count * ng, "gradients", False), MEM_HOST, float(dt),
            ptr(flux_out, count * nf, "flux_out"), ptr(isv_out, count * ni, "isv_out"),
            ptr(ct_out, count * nf * ng, "ct_out"), MEM_HOST, ctypes.byref(stats),
        )
        check(rc, "dxm_integrate_range")
        return IntegrationStats.from_c(stats)

    def read_state_into(self, key, out, gen=1):
        """Fetch one state field (``(n, dim)`` AoS) into a caller-owned array (lazy D2H of internal
        state: only ``advance()`` / ``project_on`` need it, ``quadrature_map.py:350-360``)."""
        self._require_handle()
        dim = self.variables[key] if key != "Ct" else sum(a * b for a, b in self.tangent_blocks.values())
        if out.dtype != np.float64 or not out.flags.c_contiguous or out.size != self._n * dim:
            raise ValueError(f"out must be C-contiguous float64 with {self._n * dim} entries")
        check(_lib.load().dxm_get_state(self._h, gen, key.encode(), out.ctypes.data_as(ctypes.c_void_p), MEM_HOST),
              "dxm_get_state")

    # ---- device-resident extensions ------------------------------------------------------------
    def integrate_resident(self, dt=0, wait=True):
        """Run the update on gradients already written into ``gradient_buffer()``; results stay in
        the SoA device buffers (``device_view``).  ``wait=False`` returns without synchronising."""
        if self._h is None:
            self._require_handle()
        if wait:
            rc = self._c_integrate(self._h, None, MEM_RESIDENT, dt, None, None, None, MEM_RESIDENT, self._c_stats_ref)
            self._finish(rc, self._c_stats)
            return self.last_stats
        check(self._c_integrate(self._h, None, MEM_RESIDENT, dt, None, None, None, MEM_RESIDENT, None), "dxm_integrate")
        return None

    def fetch_stats(self):
        self._require_handle()
        stats = Stats()
        check(_lib.load().dxm_last_stats(self._h, ctypes.byref(stats)), "dxm_last_stats")
        self.last_stats = IntegrationStats.from_c(stats)
        return self.last_stats

    def device_view(self, field, gen=1):
        """Zero-copy ``torch`` view (shape ``(dim, n)``, SoA) of a device-resident field via DLPack.
        Views of generation buffers are invalidated by ``data_manager.update()``.  ``"Ct"`` of a small-strain
        behaviour has 21 rows (packed symmetric storage, see ``device_tangent``)."""
        self._require_handle()
        import torch

        lib = _lib.load()
        mt = ctypes.c_void_p()
        check(lib.dxm_export_dlpack(self._h, gen, field.encode(), ctypes.byref(mt)), "dxm_export_dlpack")
        new = ctypes.pythonapi.PyCapsule_New
        new.restype = ctypes.py_object
        new.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]
        capsule = new(mt, b"dltensor", None)
        return torch.from_dlpack(capsule)

    def device_tangent(self, points=None):
        """The resident tangent as a full ``(nflux*ngrad, n)`` torch tensor.  The small-strain behaviours store
        their symmetric 6x6 tangent packed (21 rows, ``SYM6_PACKED``); this expands it on the device (a copy).
        ``points``: optional slice of Gauss points."""
        import torch

        ct = self.device_view("Ct")
        if points is not None:
            ct = ct[:, points]
        if ct.shape[0] == 21:
            return ct[torch.as_tensor(SYM6_PACKED, device=ct.device)]
        return ct

    def gradient_buffer(self):
        """Where a device-resident caller writes this step's gradients (SoA, shape ``(dim, n)``)."""
        return self.device_view(self.gradient_names[0], gen=1)

    def synth_gradients(self, seed, amp, k, K, start=0):
        """Fill the gradient buffer with the counter-based synthetic history (bench / tests)."""
        self._require_handle()
        recipe = 1 if self.behavior.finite_strain else 0
        check(_lib.load().dxm_synth_gradients(self._h, recipe, int(seed), float(amp), int(k), int(K), int(start)),
              "dxm_synth_gradients")

    def set_stream(self, cuda_stream):
        self._require_handle()
        check(_lib.load().dxm_set_stream(self._h, ctypes.c_void_p(cuda_stream)), "dxm_set_stream")

    def enable_timing(self, mode=1):
        """``kernel_ms`` of the statistics: 1 always, 0 never, -1 only for batches of >= 262144 points (default)."""
        self._require_handle()
        check(_lib.load().dxm_enable_timing(self._h, int(mode)), "dxm_enable_timing")

    def use_global_stats(self, on=True):
        """Statistics reduced over all ranks (``distributed.init_stats_comm`` first): one in-stream NCCL all-gather of
        the 64-byte record after the update kernel; every rank must then call ``integrate`` on this material."""
        self._require_handle()
        check(_lib.load().dxm_use_global_stats(self._h, int(on)), "dxm_use_global_stats")

    def enable_diagnostics(self, on=True):
        self._require_handle()
        check(_lib.load().dxm_enable_diagnostics(self._h, int(on)), "dxm_enable_diagnostics")

    def diagnostics(self):
        """Per-point ``(flag, n_iter, resid, fail)`` of the last integrate (parity tests)."""
        self._require_handle()
        flag = np.empty(self._n, dtype=np.uint8)
        fail = np.empty(self._n, dtype=np.uint8)
        n_iter = np.empty(self._n, dtype=np.int32)
        resid = np.empty(self._n)
        check(
            _lib.load().dxm_get_diagnostics(
                self._h, flag.ctypes.data_as(ctypes.c_void_p), n_iter.ctypes.data_as(ctypes.c_void_p),
                resid.ctypes.data_as(ctypes.c_void_p), fail.ctypes.data_as(ctypes.c_void_p)),
            "dxm_get_diagnostics",
        )
        return flag, n_iter, resid, fail


# ---- a true subclass of the reference's Material where the reference is importable ------------------------------------
_PLAIN = CUDAMaterial
_SUBCLASSES = {}


def material_subclass(base=None):
    """``CUDAMaterial`` as a subclass of the reference's ``dolfinx_materials.generic.Material`` (``generic.py:103-201``):
    ``isinstance(mat, Material)`` holds, every protocol member is overridden by the CUDA-backed one, and the base
    constructor runs (so ``mat.E``, ``mat.nu`` ... exist as on any reference material, ``generic.py:109-113``).
    ``base`` defaults to the installed reference class; the package exports this subclass as ``CUDAMaterial`` whenever
    ``dolfinx_materials`` can be imported, the plain protocol class otherwise."""
    if base is None:
        from dolfinx_materials.generic import Material as base  # the reference
    cls = _SUBCLASSES.get(base)
    if cls is None:

        class CUDAMaterial(_PLAIN, base):  # noqa: F811 - same public name on purpose
            __doc__ = _PLAIN.__doc__

            def __init__(self, behavior, jit=True, device=None, warn_on_failure=True):
                base.__init__(self, **dict(behavior.properties()))
                _PLAIN.__init__(self, behavior, jit=jit, device=device, warn_on_failure=warn_on_failure)

            def default_properties(self):
                return {}

        CUDAMaterial.__qualname__ = "CUDAMaterial"
        CUDAMaterial.__module__ = _PLAIN.__module__
        cls = _SUBCLASSES[base] = CUDAMaterial
    return cls


try:  # the reference package is importable only where dolfinx is (it imports dolfinx.common at module level)
    from dolfinx_materials.generic import Material as _ReferenceMaterial
except Exception:  # noqa: BLE001
    _ReferenceMaterial = None
if _ReferenceMaterial is not None:
    CUDAMaterial = material_subclass(_ReferenceMaterial)
