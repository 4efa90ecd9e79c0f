"""Pins the oracle to the REAL jaxmat arithmetic -- for an environment that has it.

The build container has neither ``jax`` nor ``jaxmat`` (un-vendored dependency of the reference, ``setup.cfg:20``), so
the jaxmat behaviours' parity is "unpinned" (DESIGN.md section 4).  Wherever ``jax``, ``equinox``, ``jaxmat`` and the
reference package are importable, one run of

    python tests/golden/make_golden_jaxmat.py [/path/to/dolfinx_materials/checkout]

drives the reference's own ``JAXMaterial`` (``dolfinx_materials/jaxmat.py:141-234``) over the histories below and
writes ``tests/golden/jaxmat_j2_voce.npz`` and ``tests/golden/jaxmat_fefp.npz``.  ``tests/test_golden_jaxmat.py``
(CPU: oracle, GPU: CUDAMaterial) picks the fixtures up when they exist and skips with "parity unpinned" otherwise.

Histories (the same seeded recipes as the oracle-side fixtures, ``oracle/synth.py``):
  * J2 + Voce, the parameters of ``demos/jax/elastoplasticity/plane_elastoplasticity.py:60-69``: 2000 points,
    4 proportional increments to amplitude 1.25e-2, ``data_manager.update()`` in between;
  * FeFp + Voce, ``tests/test_FeFp_jax.py:6-33`` verbatim (10 points, 19 steps), then 1000 points over 4 random
    ``F = I + s G`` increments to amplitude 3e-2.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.environ.get("DXM_GOLDEN_OUT", HERE)  # where the fixtures go (tests redirect it to a temporary directory)
sys.path.insert(0, ROOT)
if len(sys.argv) > 1:
    sys.path.insert(0, sys.argv[1])
elif os.path.isdir("/root/reference"):
    sys.path.insert(0, "/root/reference")

import jax  # noqa: E402

jax.config.update("jax_enable_x64", True)  # the reference path is float64

try:  # jaxmat.py:10 imports dolfinx.common.Timer; a no-op stand-in is enough where dolfinx is absent
    import dolfinx.common  # noqa: F401
except Exception:  # noqa: BLE001
    dolfinx = types.ModuleType("dolfinx")
    common = types.ModuleType("dolfinx.common")

    class Timer:
        def __init__(self, *a, **k):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    common.Timer = Timer
    dolfinx.common = common
    sys.modules["dolfinx"] = dolfinx
    sys.modules["dolfinx.common"] = common

import jaxmat  # noqa: E402
import jaxmat.materials as jm  # noqa: E402
from dolfinx_materials.jaxmat import JAXMaterial  # noqa: E402  (the reference)

from oracle import synth  # noqa: E402


def drive(material, grads):
    """``integrate`` + ``data_manager.update()`` per increment, exactly the loop of tests/test_FeFp_jax.py:27-33.
    Returns per-increment flux, isv (as returned), Ct and the final-state dict entries."""
    out = {"flux": [], "isv": [], "Ct": []}
    state = {}
    for g in grads:
        flux, isv, Ct = material.integrate(g, 0)
        out["flux"].append(np.asarray(flux, dtype=np.float64))
        out["isv"].append(np.asarray(isv, dtype=np.float64))
        out["Ct"].append(np.asarray(Ct, dtype=np.float64))
        for key, val in material.get_final_state_dict().items():
            state.setdefault("state_" + key, []).append(np.asarray(val, dtype=np.float64).reshape(len(g), -1))
        material.data_manager.update()
    res = {k: np.stack(v) for k, v in out.items()}
    res.update({k: np.stack(v) for k, v in state.items()})
    res["gradients"] = np.stack([np.asarray(g) for g in grads])
    res["isv_names"] = np.array(material.internal_state_variable_names)
    return res


def versions():
    import equinox

    return np.array([f"jax {jax.__version__}", f"equinox {equinox.__version__}",
                     f"jaxmat {getattr(jaxmat, '__version__', 'unknown')}"])


def main():
    # ---- J2 + Voce ------------------------------------------------------------------------------------------------
    props = dict(E=70e3, nu=0.3, sig0=350.0, sigu=500.0, b=1e3)
    behavior = jm.vonMisesIsotropicHardening(
        elasticity=jm.LinearElasticIsotropic(E=props["E"], nu=props["nu"]),
        yield_stress=jm.VoceHardening(sig0=props["sig0"], sigu=props["sigu"], b=props["b"]))
    n, K = 2000, 4
    material = JAXMaterial(behavior)
    material.set_data_manager(n)
    res = drive(material, [synth.strain(n, 0, 1.25e-2, k, K) for k in range(1, K + 1)])
    np.savez_compressed(os.path.join(OUT, "jaxmat_j2_voce.npz"), versions=versions(),
                        props=np.array(list(props.items()), dtype=object), **res)
    print("jaxmat_j2_voce.npz:", {k: v.shape for k, v in res.items()})

    # ---- FeFp + Voce ----------------------------------------------------------------------------------------------
    props = dict(E=70e3, nu=0.3, sig0=500.0, sigu=750.0, b=1000.0)

    def make():
        return JAXMaterial(jm.FeFpJ2Plasticity(
            elasticity=jm.LinearElasticIsotropic(E=props["E"], nu=props["nu"]),
            yield_stress=jm.VoceHardening(sig0=props["sig0"], sigu=props["sigu"], b=props["b"])))

    nb, eps, nsteps = 10, 2e-2, 20
    grads = []
    for t in np.linspace(0, 1.0, nsteps)[1:]:
        F = np.zeros((nb, 9))
        F[:, 0] = 1 + eps * t
        F[:, [1, 2]] = 1 - eps / 2 * t
        grads.append(F)
    m1 = make()
    m1.set_data_manager(nb)
    script = drive(m1, grads)
    n, K = 1000, 4
    m2 = make()
    m2.set_data_manager(n)
    rand = drive(m2, [synth.defgrad(n, 0, 3e-2, k, K) for k in range(1, K + 1)])
    np.savez_compressed(os.path.join(OUT, "jaxmat_fefp.npz"), versions=versions(),
                        props=np.array(list(props.items()), dtype=object),
                        **{"script_" + k: v for k, v in script.items()}, **{"random_" + k: v for k, v in rand.items()})
    print("jaxmat_fefp.npz:", {k: v.shape for k, v in rand.items()})


if __name__ == "__main__":
    main()
