"""dolfinx_materials_b200 -- B200-native (sm_100a) batched constitutive update behind the
``dolfinx_materials`` ``Material`` protocol.

The package is a drop-in for ONE path of bleyerj/dolfinx_materials: the per-Newton-iteration
``Material.integrate(gradients, dt)`` call that ``QuadratureMap.update`` issues over every Gauss
point (reference ``dolfinx_materials/quadrature_map.py:320-321``, ``generic.py:176-189``,
``jaxmat.py:208-234``).  Everything numerical runs in hand-written fp64 CUDA kernels loaded from
``lib/libdxm_cuda.so`` through ctypes (C ABI: ``include/dxm.h``).  There is no CPU fallback: importing
works anywhere, but creating a material's data manager without the library or without a B200 raises.
"""

__version__ = "0.1.0"


class PerformanceWarning(UserWarning):
    """Same role as ``dolfinx_materials.PerformanceWarning`` (reference ``__init__.py:12-15``);
    raised when Gauss points fail their local solve (MGIS convention, ``mfront.py:269-272``)."""


from .behaviors import (  # noqa: E402
    ElasticBehavior,
    FeFpJ2Plasticity,
    GeneralIsotropicHardening,
    Hosford,
    LinearElasticIsotropic,
    LinearHardening,
    TabulatedHardening,
    VoceHardening,
    vonMisesIsotropicHardening,
)
from .material import CUDAMaterial, IntegrationStats  # noqa: E402

__all__ = [
    "CUDAMaterial",
    "IntegrationStats",
    "PerformanceWarning",
    "LinearElasticIsotropic",
    "LinearHardening",
    "TabulatedHardening",
    "VoceHardening",
    "ElasticBehavior",
    "vonMisesIsotropicHardening",
    "FeFpJ2Plasticity",
    "GeneralIsotropicHardening",
    "Hosford",
]
