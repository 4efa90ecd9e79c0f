"""displacement vector (host) -> gradients (device) -> constitutive update, one B200.
P2 tets x 4 quadrature points (the cfg5 element), FeFp behaviour."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dolfinx_materials_b200 as jm
from dolfinx_materials_b200.fe import GradientEvaluator
from oracle import fe_gradient as fg

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 48
coords, gd, ud, nodes = fg.box_tets(nx, nx, nx, 2)
dphi = fg.tet_dphi(fg.TET_QP_DEG2, 2)
n = len(gd) * 4
mat = jm.CUDAMaterial(jm.FeFpJ2Plasticity(elasticity=jm.LinearElasticIsotropic(E=70e3, nu=0.3),
                                          yield_stress=jm.VoceHardening(sig0=500.0, sigu=750.0, b=1000.0)))
mat.set_data_manager(n)
ge = GradientEvaluator(mat, coords, gd, ud, dphi)
x, y, z = nodes.T
u = (0.02 * np.stack([x * y + 0.5 * z * z, -2 * y * z + 0.3 * x * x, 0.7 * x * z - 0.4 * y * y], axis=1)).ravel()
for _ in range(3):
    ge.eval(u); mat.integrate_resident()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
t_eval, t_tot = [], []
for _ in range(10):
    t0 = time.perf_counter(); ge.eval(u); torch.cuda.synchronize(); t1 = time.perf_counter()
    s = mat.integrate_resident(); t2 = time.perf_counter()
    t_eval.append(t1 - t0); t_tot.append(t2 - t0)
te, tt = sorted(t_eval)[3], sorted(t_tot)[3]
t0 = time.perf_counter(); g = fg.evaluate(coords, gd, ud, u, dphi, 1, 3); th = time.perf_counter() - t0
assert np.array_equal(mat.device_view("F", gen=1).cpu().numpy().T, g)
out = dict(cells=len(gd), points=n, dofs=len(nodes) * 3, h2d_bytes_u=u.nbytes, h2d_bytes_gradients_avoided=n * 72,
           eval_ms=te * 1e3, eval_gps=n / te, eval_plus_update_ms=tt * 1e3, update_kernel_ms=s.kernel_ms,
           numpy_oracle_eval_ms=th * 1e3, plastic=s.n_plastic / n)
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True); json.dump(out, open("gpurun_out/fe_gradient.json", "w"), indent=1)
