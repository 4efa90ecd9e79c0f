"""Formulation pin of the jaxmat behaviours (CPU only).

``oracle/jaxmat_form.py`` restates jaxmat's OWN formulation of the two behaviours on the hot path -- branch-free
Fischer-Burmeister local problems (scalar for ``vonMisesIsotropicHardening``, seven unknowns for ``FeFpJ2Plasticity``),
solved for every point, tangent = exact derivative of the converged stress map (what ``jax.jacfwd`` + implicit
differentiation give the reference, ``dolfinx_materials/jaxmat.py:147-164``) -- and shares no code with the canonical
oracle the CUDA kernels mirror (trial-state branch, reduced scalar / 2x2 solves, closed-form tangents).  The two must
agree at the north star's bar over whole histories: identical active sets, stress / state / tangent within rtol 1e-10.
This does not make parity green -- only vectors produced by jaxmat itself can -- but it narrows "unpinned" to
"formulation-pinned".  The last test runs the real-jaxmat fixture generator end to end against stand-in modules
(tests/fake_jaxmat) so that the hook and the tests that consume its fixtures cannot rot."""
import os
import subprocess
import sys

import numpy as np
import pytest
from golden_check import close

from oracle import fefp, synth
from oracle import jaxmat_form as jf
from oracle import small_strain as ss

HERE = os.path.dirname(os.path.abspath(__file__))
J2 = dict(E=70e3, nu=0.3, sig0=350.0, sigu=500.0, b=1e3)  # plane_elastoplasticity.py:60-69
FE = dict(E=70e3, nu=0.3, sig0=500.0, sigu=750.0, b=1000.0)  # tests/test_FeFp_jax.py:7-15


def _props(g):
    return dict(zip([str(k) for k in g["props_keys"]], [float(v) for v in g["props_vals"]]))


def test_j2_voce_golden_history_in_jaxmat_form():
    """Every point of tests/golden/j2_voce_history.npz (the fixture the canonical oracle and the kernels are held to)."""
    g = np.load(os.path.join(HERE, "golden", "j2_voce_history.npz"))
    props, n = _props(g), g["eps1"].shape[0]
    st, k = ss.zero_state(n), 1
    while f"eps{k}" in g:
        ref = ss.integrate(g[f"eps{k}"], st, props)
        out = jf.j2_integrate(g[f"eps{k}"], st, props)
        assert np.array_equal(out["flag"], ref["flag"]), f"active set differs at increment {k}"
        for got, want in ((out["stress"], g[f"flux{k}"]), (out["p"], g[f"isv{k}"][:, 0]), (out["epsp"], g[f"isv{k}"][:, 1:]),
                          (out["Ct"], g[f"Ct{k}"])):
            close(got, want)
        st = ss.advance(ref)
        k += 1
    assert k == 5 and ref["flag"].any() and not ref["flag"].all()


def test_fefp_golden_history_in_jaxmat_form():
    g = np.load(os.path.join(HERE, "golden", "fefp_history.npz"))
    props, n = _props(g), g["F1"].shape[0]
    st, k = fefp.virgin_state(n), 1
    while f"F{k}" in g:
        ref = fefp.integrate(g[f"F{k}"], st, props)
        out = jf.fefp_integrate(g[f"F{k}"], st, props)
        assert np.array_equal(out["flag"], ref["flag"]), f"active set differs at increment {k}"
        for got, want in ((out["PK1"], g[f"flux{k}"]), (out["p"], g[f"isv{k}"][:, 0]), (out["be_bar"], g[f"isv{k}"][:, 1:]),
                          (out["Ct"], g[f"Ct{k}"])):
            close(got, want)
        st = fefp.advance(ref)
        k += 1
    assert k == 4 and ref["flag"].any()


def test_seeded_histories_in_jaxmat_form():
    """BASELINE configs 2 and 3 at oracle size: 4000 / 2000 points, four increments each, state carried."""
    n, K = 4000, 4
    st = ss.zero_state(n)
    for k in range(1, K + 1):
        eps = synth.strain(n, 0, 1.25e-2, k, K)
        ref, out = ss.integrate(eps, st, J2), jf.j2_integrate(eps, st, J2)
        assert np.array_equal(out["flag"], ref["flag"])
        for f in ("stress", "p", "epsp", "Ct"):
            close(out[f], ref[f], f)
        st = ss.advance(ref)
    assert 0.5 < ref["flag"].mean() < 0.8
    n = 2000
    st = fefp.virgin_state(n)
    for k in range(1, K + 1):
        F = synth.defgrad(n, 0, 6e-2, k, K)
        ref, out = fefp.integrate(F, st, FE), jf.fefp_integrate(F, st, FE)
        assert np.array_equal(out["flag"], ref["flag"])
        for f in ("PK1", "p", "be_bar", "Ct"):
            close(out[f], ref[f], f)
        st = fefp.advance(ref)
    assert ref["flag"].mean() > 0.7 and ref["fail"].sum() == 0


def test_reference_test_script_in_jaxmat_form():
    """tests/test_FeFp_jax.py:6-33 verbatim (10 points, 19 uniaxial steps): the known answers of SURVEY 8(c)(4)."""
    nb, eps = 10, 2e-2
    st = fefp.virgin_state(nb)
    for t in np.linspace(0, 1.0, 20)[1:]:
        F = np.zeros((nb, 9))
        F[:, 0] = 1 + eps * t
        F[:, [1, 2]] = 1 - eps / 2 * t
        ref, out = fefp.integrate(F, st, FE), jf.fefp_integrate(F, st, FE)
        assert np.array_equal(out["flag"], ref["flag"])
        for f in ("PK1", "p", "be_bar", "Ct"):
            close(out[f], ref[f], f)
        st = fefp.advance(ref)
    assert abs(out["p"][0] - 1.076097e-2) < 1e-8 and abs(out["PK1"][0, 0] - 473.1527) < 1e-3


def test_generator_runs_end_to_end_against_stand_in_jaxmat(tmp_path):
    """make_golden_jaxmat.py, unmodified, in an environment where `jax`, `equinox`, `jaxmat` and
    `dolfinx_materials.jaxmat` resolve to the stand-ins of tests/fake_jaxmat: it must write fixtures in its real format,
    and the consuming comparison (tests/test_golden_jaxmat.py::check_history) must hold the canonical oracle to them."""
    import test_golden_jaxmat as consumer

    fake = os.path.join(HERE, "fake_jaxmat")
    env = dict(os.environ, DXM_GOLDEN_OUT=str(tmp_path), PYTHONPATH=fake + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, os.path.join(HERE, "golden", "make_golden_jaxmat.py"), fake], env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    fix = dict(np.load(tmp_path / "jaxmat_j2_voce.npz", allow_pickle=True))
    assert "standin" in str(fix["versions"][0]) and fix["gradients"].shape == (4, 2000, 6)
    ref = consumer.check_history(fix, "", ss.integrate, ss.zero_state(2000), consumer.J2_PROPS,
                                 ("stress", consumer.isv_names(fix, "", None)))
    assert ref["flag"].any()
    fix = dict(np.load(tmp_path / "jaxmat_fefp.npz", allow_pickle=True))
    assert fix["script_gradients"].shape == (19, 10, 9) and fix["random_gradients"].shape == (4, 1000, 9)
    for prefix in ("script_", "random_"):
        n = fix[prefix + "gradients"].shape[1]
        ref = consumer.check_history(fix, prefix, fefp.integrate, fefp.virgin_state(n), consumer.FEFP_PROPS,
                                     ("PK1", consumer.isv_names(fix, prefix, None)))
        assert ref["flag"].any()
