"""CPU oracle for the batched constitutive update -- TEST INFRASTRUCTURE ONLY.

This package restates, in plain numpy fp64, the algorithm behind the reference's
``Material.integrate(gradients, dt)`` hot path (reference files, all relative to the
upstream tree: ``dolfinx_materials/generic.py:176-189`` for the protocol,
``dolfinx_materials/jaxmat.py:141-234`` for the batched jaxmat back-end,
``dolfinx_materials/python_materials/elasticity.py:5-24`` for linear elasticity and
``tests/mfront/IsotropicLinearHardeningPlasticity.mfront:49-77`` for the closed-form J2
radial return and its consistent tangent).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product package
(``dolfinx_materials_b200``) never does: it fails loudly when its CUDA library is missing.

PARITY STATUS
-------------
* elastic update and the ``Material``/``DataManager`` state machinery: PINNED against the
  reference's own ``generic.Material`` + ``LinearElasticIsotropic`` executed in the build
  container (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).
* J2 (linear hardening): pinned against the closed form of the in-tree MFront source
  (formula restatement) and the analytic plane-strain limit asserted by the reference's
  ``tests/mfront/test_elastoplasticity.py:31-36``.
* J2 + Voce and FeFp arithmetic: **parity unpinned** -- it lives in the third-party,
  un-vendored ``jaxmat`` package (``setup.cfg:20``: ``jaxmat>=0.0.1``, no lock file), which is
  not installable here, and the reference's only test on that path
  (``tests/test_FeFp_jax.py``) asserts nothing.  The restatement follows the published
  algorithm and is checked by self-consistency (finite-difference
  tangents, yield consistency, det(be_bar)=1) and against an independent complex-arithmetic
  statement of jaxmat's Fischer-Burmeister residual systems whose solution is differentiated
  exactly by the complex-step method -- the definition of the reference's jacfwd tangent
  (``tests/test_oracle_exact_derivative.py``: stress to 1e-11, tangent to 1e-10).

* Hosford (``oracle/hosford.py``, plain C): **parity unpinned** against MFront (TFEL/MGIS absent, generated code not
  in the tree); restates ``demos/multimaterials/IsotropicPlasticHosfordFlowLinear.mfront`` and is checked against an
  independent statement of the implicit system, finite-difference tangents, a = 2 == the J2 oracle and the
  criterion's known yield points (``tests/test_oracle_hosford.py``).

* distance to the EXACT solution (``oracle/jaxmat_form_mp.py``, ``oracle/hosford_mp.py``): jaxmat's branch-free
  systems and the seven-unknown system of MFront's ``Implicit`` DSL solved point by point in 40-digit arithmetic
  (``mpmath``), tangents as central differences of the solution map: the canonical arithmetic is within 2e-12 (stress),
  the local Newton tolerance (plastic multiplier) and 1e-11 (tangent) of it
  (``tests/test_oracle_jaxmat_form_mp.py``, ``tests/test_oracle_hosford_mp.py``).  That pins the mathematics, not the
  dependencies' own floating-point output.

Canonical arithmetic
--------------------
Every function is written component-wise with an explicit operation order and uses only
IEEE-754 correctly rounded operations (+, -, *, /, sqrt, rint, ldexp); ``exp`` is the
hand-written :func:`oracle.canon.exp_c`.  The CUDA kernels are compiled with
``-fmad=false`` and follow the same order, so kernel and oracle agree **bit for bit**
(flags, local iteration counts, stress, state and tangent).
"""

from . import canon, fefp, small_strain, synth  # noqa: F401
