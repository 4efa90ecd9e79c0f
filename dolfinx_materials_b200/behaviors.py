"""Behaviour descriptors with the constructor signatures of the ``jaxmat.materials`` objects the
reference wraps (call sites: ``demos/jax/elastoplasticity/plane_elastoplasticity.py:60-73``,
``demos/jax/finite_strain_elastoplasticity/finite_strain_elastoplasticity.py:158-169``,
``tests/test_FeFp_jax.py:7-19``).  They carry parameters only -- the arithmetic lives in the CUDA
kernels -- so that a reference script switches by replacing

    import jaxmat.materials as jm ; material = JAXMaterial(behavior)
by
    import dolfinx_materials_b200 as jm ; material = jm.CUDAMaterial(behavior)
"""

from dataclasses import dataclass, field
from typing import Any

# behaviour ids of include/dxm.h
DXM_ELASTIC, DXM_J2_LINEAR, DXM_J2_VOCE, DXM_FEFP_VOCE, DXM_J2_TABLE, DXM_HOSFORD_LINEAR = 0, 1, 2, 3, 4, 5


@dataclass
class LinearElasticIsotropic:
    """Isotropic linear elasticity (``python_materials/elasticity.py:5-19``)."""

    E: Any
    nu: Any


@dataclass
class LinearHardening:
    """sigma_Y(p) = sig0 + H p  (``tests/mfront/IsotropicLinearHardeningPlasticity.mfront:11``)."""

    sig0: Any
    H: Any = 0.0


@dataclass
class VoceHardening:
    """sigma_Y(p) = sig0 + (sigu - sig0) (1 - exp(-b p)) [+ H p]  (``tests/test_FeFp_jax.py:14-15``)."""

    sig0: Any
    sigu: Any
    b: Any
    H: Any = 0.0


@dataclass
class TabulatedHardening:
    """Piecewise-linear isotropic hardening through ``(p[k], sig[k])`` (``p[0] = 0``, increasing, at most 64 points),
    continued with the last slope.  ``vonMisesIsotropicHardening`` accepts an arbitrary ``yield_stress`` callable in
    jaxmat (old demo ``demos/jax/elastoplasticity/_plane_stress_elastoplasticity.py:38-44``); a Python callable cannot be
    compiled into the CUDA kernels, so ``from_callable`` samples it (denser near ``p = 0`` where hardening curves bend)."""

    p: Any
    sig: Any

    @classmethod
    def from_callable(cls, yield_stress, p_max, n=64):
        import numpy as np

        p = np.concatenate([[0.0], np.geomspace(p_max * 1e-5, p_max, n - 1)])
        return cls(p=p, sig=np.array([float(yield_stress(x)) for x in p]))


def _hardening_props(h):
    if isinstance(h, TabulatedHardening):
        return {}
    if isinstance(h, LinearHardening):
        return {"sig0": h.sig0, "H": h.H}
    if isinstance(h, VoceHardening):
        return {"sig0": h.sig0, "sigu": h.sigu, "b": h.b, "H": h.H}
    raise TypeError(
        "yield_stress must be a LinearHardening, VoceHardening or TabulatedHardening descriptor; arbitrary Python "
        "callables cannot be compiled into the CUDA kernels (sample them with TabulatedHardening.from_callable)"
    )


@dataclass
class _Behavior:
    elasticity: LinearElasticIsotropic
    kind: int = field(init=False, default=DXM_ELASTIC)
    finite_strain: bool = field(init=False, default=False)

    def properties(self):
        return {"E": self.elasticity.E, "nu": self.elasticity.nu}


@dataclass
class ElasticBehavior(_Behavior):
    """Small-strain linear elasticity."""


@dataclass
class vonMisesIsotropicHardening(_Behavior):
    """Small-strain J2 plasticity with isotropic hardening."""

    yield_stress: Any = None

    def __post_init__(self):
        if isinstance(self.yield_stress, TabulatedHardening):
            self.kind = DXM_J2_TABLE
        else:
            self.kind = DXM_J2_LINEAR if isinstance(self.yield_stress, LinearHardening) else DXM_J2_VOCE

    def properties(self):
        return {**super().properties(), **_hardening_props(self.yield_stress)}


@dataclass
class Hosford:
    """Hosford equivalent stress ``(1/2 (|s1-s2|^a + |s2-s3|^a + |s3-s1|^a))^(1/a)`` with an even integer exponent
    (``a = 2``: von Mises; ``criterion : "Hosford" {a: 10}`` in
    ``demos/multimaterials/IsotropicPlasticHosfordFlowLinear.mfront:21``)."""

    a: int = 10


@dataclass
class GeneralIsotropicHardening(_Behavior):
    """Small-strain associated plasticity with a non-quadratic isotropic criterion and isotropic hardening.  With
    ``LinearHardening`` (``R0 + H p``) it is the MFront behaviour of the matrix phase of
    ``demos/multimaterials/multimaterials.py:245-254`` (``young_modulus, poisson_ratio, R0, hardening_slope`` =
    ``E, nu, sig0, H`` here); with ``VoceHardening`` it is the jaxmat behaviour of that name as the old demo calls it,
    ``GeneralIsotropicHardening(elastic_model, yield_stress, ...)`` (``_plane_stress_elastoplasticity.py:17,45``)."""

    yield_stress: Any = None
    equivalent_stress: Any = field(default_factory=Hosford)

    def __post_init__(self):
        if not isinstance(self.yield_stress, (LinearHardening, VoceHardening)):
            raise TypeError("GeneralIsotropicHardening: yield_stress must be a LinearHardening or VoceHardening descriptor")
        if not isinstance(self.equivalent_stress, Hosford):
            raise TypeError("GeneralIsotropicHardening: equivalent_stress must be a Hosford descriptor")
        a = self.equivalent_stress.a
        if int(a) != a or a < 2 or a > 64 or int(a) % 2:
            raise ValueError("Hosford exponent: an even integer in [2, 64]")
        self.kind = DXM_HOSFORD_LINEAR

    def properties(self):
        return {**super().properties(), **_hardening_props(self.yield_stress), "a": int(self.equivalent_stress.a)}


@dataclass
class FeFpJ2Plasticity(_Behavior):
    """Finite-strain multiplicative (Fe.Fp) J2 plasticity, state ``be_bar`` + ``p``."""

    yield_stress: Any = None

    def __post_init__(self):
        self.kind = DXM_FEFP_VOCE
        self.finite_strain = True

    def properties(self):
        return {**super().properties(), **_hardening_props(self.yield_stress)}
