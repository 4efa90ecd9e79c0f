"""Config-5 surrogate (scripts/newton_bar.py): the device-resident Newton loop -- gradients, FeFp update, fused
assembly, Krylov solve -- reproduces the same loop run with the CPU oracles and a direct sparse solve."""
import os
import sys

import numpy as np
import pytest

from oracle import fe_forms as ff
from oracle import fe_gradient as fg
from oracle import fefp

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
import newton_bar as nb  # noqa: E402

pytestmark = pytest.mark.gpu

PROPS = dict(E=70e3, nu=0.3, sig0=500.0, sigu=750.0, b=1000.0)


def oracle_newton(nx, ny, nz, steps, strain, L=10.0, W=1.0):
    from scipy.sparse.linalg import spsolve

    nodes, gd, ud, _ = nb.bar_mesh(nx, ny, nz, L, W)
    dphi = nb.p2_tet_dphi(nb.QP_DEG2)
    bc, top = nb.boundary_conditions(nodes, L)
    st = fefp.virgin_state(len(gd) * 4)
    u = np.zeros(3 * len(nodes))
    iters = []
    for k in range(1, steps + 1):
        lift = np.zeros_like(u)
        lift[top] = -(strain * L / steps)
        r0 = None
        for it in range(25):
            out = fefp.integrate(fg.evaluate(nodes, gd, ud, u, dphi, 1, 3), st, PROPS)
            fe, ke = ff.element_forms(nodes, gd, ud, dphi, nb.W_DEG2, out["PK1"], out["Ct"], 1, 3)
            b, A = ff.assemble(ud, fe, ke, len(nodes), 3, bc=bc, lift=lift if it == 0 else None)
            rn = np.linalg.norm(b)
            r0 = rn if r0 is None else r0
            if it > 0 and (rn <= 1e-8 or rn <= 1e-8 * r0):
                break
            u -= spsolve(A.tocsc(), b)
        iters.append(it)
        st = fefp.advance(out)
    return u, out, iters


def test_device_newton_matches_oracle_newton(jm):
    nx, ny, nz, steps, strain = 5, 1, 1, 3, 0.012
    u_ref, out_ref, iters_ref = oracle_newton(nx, ny, nz, steps, strain)
    assert out_ref["flag"].any() and not out_ref["flag"].all()
    u, mat, info, hist = nb.run_gpu(nx, ny, nz, steps=steps, strain=strain, ksp_rtol=1e-11, verbose=False)
    assert info["newton_iterations"] == sum(iters_ref)
    assert np.allclose(u, u_ref, rtol=1e-7, atol=1e-9 * np.abs(u_ref).max())
    p = mat.get_initial_state_dict()["p"].ravel()  # s0 after the last update() == converged state
    assert np.allclose(p, out_ref["p"], rtol=1e-6, atol=1e-12)
    assert np.array_equal(p > 0, out_ref["p"] > 0)
