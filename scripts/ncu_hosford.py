"""Workload for the ncu capture of the Hosford kernels (split launch: light + heavy; DXM_HOS_SPLIT=0: fused): 4e6 points, the demo's parameters (a = 10), second increment of
a two-increment history (about half the points plastic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dolfinx_materials_b200 as jm

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000
m = jm.CUDAMaterial(jm.GeneralIsotropicHardening(elasticity=jm.LinearElasticIsotropic(E=70e3, nu=0.3),
                                                 yield_stress=jm.LinearHardening(sig0=200.0, H=10.0)))
m.set_data_manager(n)
m.synth_gradients(0, 1.25e-2, 1, 2); m.integrate_resident(); m.data_manager.update()
m.synth_gradients(0, 1.25e-2, 2, 2)
for _ in range(3):
    s = m.integrate_resident()
print("n", n, "plastic", s.n_plastic / n, "max_iter", s.max_iter, "fail", s.n_fail, "kernel_ms", s.kernel_ms, "gps", n / s.kernel_ms * 1e3)
