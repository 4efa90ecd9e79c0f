"""Small-batch launches for an ncu launch list (pure kernel durations at n = 1, 1e3, 1e4, 1e5)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dolfinx_materials_b200 as jm
for n in (1, 1000, 10_000, 100_000):
    m = jm.CUDAMaterial(jm.vonMisesIsotropicHardening(elasticity=jm.LinearElasticIsotropic(E=70e3, nu=0.3),
                                                       yield_stress=jm.VoceHardening(sig0=350.0, sigu=500.0, b=1e3)))
    m.set_data_manager(n)
    m.synth_gradients(0, 1.25e-2, 1, 1)
    for _ in range(4):
        m.integrate_resident()
