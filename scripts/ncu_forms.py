"""Workload for the ncu captures of the FE-side kernels: P2 tets, gradients -> FeFp update -> fused assembly."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import newton_bar as nb
import dolfinx_materials_b200 as jm
from dolfinx_materials_b200.fe import AssembledSystem, ElementForms, GradientEvaluator

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
coords, gd, ud, nodes = nb.bar_mesh(nx, 24, 24)
mat = jm.CUDAMaterial(jm.FeFpJ2Plasticity(elasticity=jm.LinearElasticIsotropic(E=70e3, nu=0.3),
                                          yield_stress=jm.VoceHardening(sig0=500.0, sigu=750.0, b=1000.0)))
mat.set_data_manager(len(gd) * 4)
ge = GradientEvaluator(mat, coords, gd, ud, nb.p2_tet_dphi(nb.QP_DEG2))
forms = ElementForms(ge, nb.W_DEG2)
rowptr, colidx = nb.sparsity(ud, len(nodes))
bc, top = nb.boundary_conditions(nodes, 10.0)
system = AssembledSystem(forms, rowptr, colidx, bc=bc)
x, y, z = nodes.T
u = (0.004 * np.stack([x, -0.3 * y, -0.3 * z], axis=1)).ravel()
for _ in range(3):
    ge.eval(u); s = mat.integrate_resident(); system.assemble()
print("cells", len(gd), "plastic", s.n_plastic / (len(gd) * 4), "kernel_ms", s.kernel_ms)
