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


def identify_hardening(yield_stress, p_max=10.0):
    """Recognise the law behind a ``yield_stress(p)`` callable -- what ``jaxmat`` behaviours take
    (``tests/test_FeFp_jax.py:14-19``, old demo ``_plane_stress_elastoplasticity.py:38-44``) -- as an instance of the
    family the kernels implement, ``sigma_Y(p) = sig0 + H p + (sigu - sig0) (1 - exp(-b p))``, and return the matching
    :class:`LinearHardening` / :class:`VoceHardening` descriptor, or ``None`` when the callable is something else.

    The callable is only *probed* (called with Python floats on ``[0, p_max]``; anything ``float()`` accepts may come
    back, numpy or jax scalars alike) -- it never runs per Gauss point.  Parameters are found by variable projection
    (``b`` on a log scale, ``H`` and ``sigu - sig0`` linearly), snapped to the shortest decimal numbers that reproduce
    the samples to rounding (the callable's own constants, when it is of the family), and accepted only if the law
    then matches the callable to 1e-11 of its magnitude on an independent set of points."""
    import numpy as np

    def f(x):
        return float(np.asarray(yield_stress(float(x))).reshape(()))

    try:
        sig0 = f(0.0)
        fit_p = np.concatenate([np.geomspace(1e-9 * p_max, p_max, 160), np.linspace(p_max / 40, p_max, 40)])
        chk_p = np.geomspace(3.3e-10 * p_max, 0.97 * p_max, 97)
        fit_g = np.array([f(x) for x in fit_p]) - sig0
        chk_g = np.array([f(x) for x in chk_p]) - sig0
    except Exception:  # noqa: BLE001 - a callable that cannot be probed with floats is simply not recognised
        return None
    if not (np.isfinite(sig0) and np.all(np.isfinite(fit_g)) and np.all(np.isfinite(chk_g))):
        return None
    scale = max(abs(sig0), np.abs(fit_g).max(), 1e-300)
    all_p, all_g = np.concatenate([fit_p, chk_p]), np.concatenate([fit_g, chk_g])

    def model(H, dsu, b, x):
        return H * x + dsu * (1.0 - np.exp(-(b * x)))

    def err(H, dsu, b, x=all_p, g=all_g):
        return np.abs(model(H, dsu, b, x) - g).max() / scale

    def snap(vals, i, offset=0.0, tol=1e-14):
        """Shortest decimal value of parameter i (+ offset) that keeps the law on the samples to rounding."""
        if err(*vals) > tol:
            return vals
        for digits in range(1, 16):
            v = vals[i] + offset
            if v == 0.0:
                break
            r = float(f"{v:.{digits - 1}e}") - offset
            trial = list(vals)
            trial[i] = r
            if err(*trial) <= tol:
                return trial
        return vals

    # no saturation term: sig0 + H p
    H = float(fit_g[-1] / fit_p[-1])
    if err(H, 0.0, 0.0) <= 1e-13:
        H = snap([H, 0.0, 0.0], 0)[0]
        return LinearHardening(sig0=sig0, H=H)

    def project(b):
        A = np.stack([fit_p, 1.0 - np.exp(-(b * fit_p))], axis=1)
        coef = np.linalg.lstsq(A / scale, fit_g / scale, rcond=None)[0]
        return coef, np.abs(A @ coef - fit_g).max() / scale

    grid = np.geomspace(1e-4 / p_max, 1e9 / p_max, 400)
    res = [project(b)[1] for b in grid]
    k = int(np.argmin(res))
    lo, hi = np.log(grid[max(k - 1, 0)]), np.log(grid[min(k + 1, len(grid) - 1)])
    for _ in range(200):  # golden-section search on log b of the projected residual
        m1, m2 = hi - 0.6180339887498949 * (hi - lo), lo + 0.6180339887498949 * (hi - lo)
        if project(np.exp(m1))[1] < project(np.exp(m2))[1]:
            hi = m2
        else:
            lo = m1
    b = float(np.exp(0.5 * (lo + hi)))
    (H, dsu), _ = project(b)
    vals = [float(H), float(dsu), b]
    for _ in range(20):  # Gauss-Newton polish of all three parameters
        H, dsu, b = vals
        e = np.exp(-(b * fit_p))
        J = np.stack([fit_p, 1.0 - e, dsu * fit_p * e], axis=1) / scale
        step = np.linalg.lstsq(J, (fit_g - model(H, dsu, b, fit_p)) / scale, rcond=None)[0]
        trial = [H + step[0], dsu + step[1], b + step[2]]
        if not np.all(np.isfinite(trial)) or trial[2] <= 0 or err(*trial, fit_p, fit_g) >= err(*vals, fit_p, fit_g):
            break
        vals = [float(v) for v in trial]
    vals = snap(vals, 2)             # b
    vals = snap(vals, 1, sig0)       # sigu = sig0 + dsu
    if abs(vals[0]) * p_max <= 1e-13 * scale:
        vals[0] = 0.0
    vals = snap(vals, 0)             # H
    if err(*vals, chk_p, chk_g) > 1e-11:
        return None
    H, dsu, b = vals
    return VoceHardening(sig0=sig0, sigu=sig0 + dsu, b=b, H=H)


def _resolve_hardening(h):
    """Descriptors pass through; a callable is recognised as a member of the kernels' hardening family or refused."""
    if isinstance(h, (LinearHardening, VoceHardening, TabulatedHardening)) or not callable(h):
        return h
    found = identify_hardening(h)
    if found is None:
        raise TypeError(
            "yield_stress: the callable is not of the form sig0 + H p + (sigu - sig0) (1 - exp(-b p)) that the CUDA "
            "kernels implement, and arbitrary Python callables cannot be compiled into them; pass a LinearHardening / "
            "VoceHardening descriptor, or sample the curve with TabulatedHardening.from_callable (small-strain J2)"
        )
    return found


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
        self.yield_stress = _resolve_hardening(self.yield_stress)
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
        self.yield_stress = _resolve_hardening(self.yield_stress)
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
        self.yield_stress = _resolve_hardening(self.yield_stress)
        if isinstance(self.yield_stress, TabulatedHardening):
            raise TypeError("FeFpJ2Plasticity: tabulated hardening is a small-strain J2 option (vonMisesIsotropicHardening)")
        self.kind = DXM_FEFP_VOCE
        self.finite_strain = True

    def properties(self):
        return {**super().properties(), **_hardening_props(self.yield_stress)}
