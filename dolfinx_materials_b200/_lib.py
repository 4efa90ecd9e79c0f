"""ctypes binding of ``lib/libdxm_cuda.so`` (C ABI declared in ``include/dxm.h``).

There is deliberately no fallback: if the shared library is missing or cannot be loaded the
first use raises ``RuntimeError``.
"""

import ctypes
import pathlib

import os

# DXM_UNFUSED=1: the A/B build with the round-1 (un-fused) arithmetic; DXM_VARIANT=<name>: an experiment build
# (lib/libdxm_cuda_<name>.so), see build.py
_VARIANT = "unfused" if os.environ.get("DXM_UNFUSED", "0") not in ("", "0") else os.environ.get("DXM_VARIANT", "")
LIB_PATH = pathlib.Path(__file__).resolve().parent / "lib" / (
    f"libdxm_cuda_{_VARIANT}.so" if _VARIANT else "libdxm_cuda.so")

MEM_HOST, MEM_DEVICE, MEM_RESIDENT = 0, 1, 2

c_double_p = ctypes.POINTER(ctypes.c_double)


class Stats(ctypes.Structure):
    _fields_ = [
        ("n_points", ctypes.c_int64),
        ("n_plastic", ctypes.c_int64),
        ("n_fail", ctypes.c_int64),
        ("max_iter", ctypes.c_int64),
        ("max_residual", ctypes.c_double),
        ("kernel_ms", ctypes.c_double),
    ]


# name -> (restype, argtypes); must list every symbol of include/dxm.h
SIGNATURES = {
    "dxm_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.POINTER(ctypes.c_void_p)]),
    "dxm_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "dxm_device_count": (ctypes.c_int, []),
    "dxm_set_stream": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "dxm_ld": (ctypes.c_int64, [ctypes.c_void_p]),
    "dxm_npoints": (ctypes.c_int64, [ctypes.c_void_p]),
    "dxm_set_property": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]),
    "dxm_set_hardening_table": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "dxm_field_dim": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p]),
    "dxm_set_state": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int]),
    "dxm_get_state": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int]),
    "dxm_device_ptr": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]),
    "dxm_export_dlpack": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]),
    "dxm_integrate": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p,
         ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(Stats)],
    ),
    "dxm_integrate_range": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_void_p,
         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(Stats)],
    ),
    "dxm_last_stats": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(Stats)]),
    "dxm_enable_timing": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "dxm_comm_unique_id": (ctypes.c_int, [ctypes.c_void_p]),
    "dxm_comm_init": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "dxm_comm_p2p_handle": (ctypes.c_int, [ctypes.c_void_p]),
    "dxm_comm_p2p_connect": (ctypes.c_int, [ctypes.c_void_p]),
    "dxm_comm_p2p_enabled": (ctypes.c_int, []),
    "dxm_comm_p2p_disable": (ctypes.c_int, []),
    "dxm_comm_size": (ctypes.c_int, []),
    "dxm_comm_rank": (ctypes.c_int, []),
    "dxm_comm_destroy": (ctypes.c_int, []),
    "dxm_use_global_stats": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "dxm_update": (ctypes.c_int, [ctypes.c_void_p]),
    "dxm_revert": (ctypes.c_int, [ctypes.c_void_p]),
    "dxm_enable_diagnostics": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "dxm_get_diagnostics": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "dxm_synth_gradients": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int64],
    ),
    "dxm_mesh_create": (
        ctypes.c_int,
        [ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
         ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)],
    ),
    "dxm_mesh_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "dxm_eval_gradient": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "dxm_mesh_set_weights": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "dxm_element_forms": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "dxm_system_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    "dxm_system_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "dxm_system_set_bc": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "dxm_system_set_lifting": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "dxm_assemble": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "dxm_system_get": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "dxm_system_defer_constraints": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "dxm_system_apply_constraints": (ctypes.c_int, [ctypes.c_void_p]),
    "dxm_system_device_ptrs": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p)]),
    "dxm_system_nnz": (ctypes.c_int64, [ctypes.c_void_p]),
    "dxm_system_solve": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                        ctypes.POINTER(ctypes.c_int), c_double_p]),
    "dxm_host_alloc": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int64]),
    "dxm_host_free": (ctypes.c_int, [ctypes.c_void_p]),
    "dxm_host_register": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64]),
    "dxm_host_unregister": (ctypes.c_int, [ctypes.c_void_p]),
    "dxm_host_mirror_sym6": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]),
    "dxm_host_gather_rows": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int]),
    "dxm_host_scatter_rows": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int]),
    "dxm_launch_count": (ctypes.c_int64, []),
    "dxm_fp64_peak": (ctypes.c_int, [ctypes.c_int, c_double_p]),
    "dxm_copy_peak": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, c_double_p]),
    "dxm_stream_peak": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int, c_double_p]),
    "dxm_last_error": (ctypes.c_char_p, []),
    "dxm_version": (ctypes.c_char_p, []),
}

_lib = None


def load():
    """Load the library once; raise loudly if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m dolfinx_materials_b200.build` "
            "(or __graft_entry__.build()). dolfinx_materials_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the header drift apart
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


class DxmError(RuntimeError):
    pass


def check(rc, what=""):
    if rc < 0:
        msg = load().dxm_last_error().decode("utf-8", "replace")
        raise DxmError(f"{what}: {msg}" if what else msg)
    return rc
