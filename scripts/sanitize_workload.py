"""Small workload that touches every kernel of the library, for compute-sanitizer (memcheck / racecheck / initcheck):
host-pipeline J2 (packed tangent, compaction on and off; DXM_HOST_MIRROR=0/1 from the environment), Hosford, FeFp (compacted), per-point properties, gradient evaluation,
element forms, assembly with constraints + lifting, Krylov solve."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import newton_bar as nb
import dolfinx_materials_b200 as jm
from dolfinx_materials_b200.fe import AssembledSystem, ElementForms, GradientEvaluator
from oracle import synth

el = jm.LinearElasticIsotropic(E=70e3, nu=0.3)
n = 3001
for beh in (jm.vonMisesIsotropicHardening(elasticity=el, yield_stress=jm.VoceHardening(sig0=350.0, sigu=500.0, b=1e3)),
            jm.vonMisesIsotropicHardening(elasticity=el, yield_stress=jm.LinearHardening(sig0=250.0, H=5e3)),
            jm.GeneralIsotropicHardening(elasticity=el, yield_stress=jm.LinearHardening(sig0=200.0, H=10.0)),
            jm.ElasticBehavior(elasticity=el)):
    m = jm.CUDAMaterial(beh); m.set_data_manager(n)
    for k in range(1, 4):
        os.environ["DXM_HOS_SPLIT"] = str(k % 2)  # Hosford: fused and tiled kernels
        flux, isv, ct = m.integrate(synth.strain(n, 0, 1.25e-2, k, 3)); m.data_manager.update()
    m.update_material_property("E", np.linspace(60e3, 80e3, n)); m.integrate(synth.strain(n, 0, 1.3e-2, 3, 3))
    m.device_tangent(); m.get_final_state_dict()
fm = jm.CUDAMaterial(jm.FeFpJ2Plasticity(elasticity=el, yield_stress=jm.VoceHardening(sig0=500.0, sigu=750.0, b=1000.0)))
fm.set_data_manager(n)
for k in range(1, 4):
    fm.integrate(synth.defgrad(n, 0, 1.2e-2 * k / 3 + 4e-3, k, 3)); fm.data_manager.update()
u, mat, info, hist = nb.run_gpu(3, 1, 1, steps=2, strain=0.012, ksp_rtol=1e-10, verbose=False)
coords, gd, ud, nodes = nb.bar_mesh(3, 1, 1)
ge = GradientEvaluator(mat, coords, gd, ud, nb.p2_tet_dphi(nb.QP_DEG2)); ge.eval(u); mat.integrate_resident()
ElementForms(ge, nb.W_DEG2).compute()
print("sanitize workload ok", info["newton_iterations"])
# round 2: the atomic-free per-node gather assembly, the asynchronous / ranged statistics paths, small-batch launches
os.environ["DXM_FE_GATHER"] = "1"
u2, mat2, info2, _ = nb.run_gpu(3, 2, 1, steps=1, strain=0.012, ksp_rtol=1e-10, verbose=False)
del os.environ["DXM_FE_GATHER"]
for nn in (1, 63, 64, 65, 5000):
    ms = jm.CUDAMaterial(jm.vonMisesIsotropicHardening(elasticity=el, yield_stress=jm.VoceHardening(sig0=350.0, sigu=500.0, b=1e3)))
    ms.set_data_manager(nn); ms.synth_gradients(0, 1.25e-2, 1, 1)
    ms.integrate_resident(); ms.integrate_resident(wait=False); ms.fetch_stats()
    ms.enable_timing(1); ms.integrate_resident()
ms.integrate_range_into(0, 2000, synth.strain(2000, 0, 1e-2, 1, 1), np.empty((2000, 6)), None, np.empty((2000, 36)))
print("sanitize workload (round 2 additions) ok", info2["newton_iterations"])
# last session of round 2: host arrays through the small-batch path (mapped page-locked staging, <= 2048 points), the
# rewritten fe_forms_kernel (fast and rolled paths: run_gpu above assembles with constraints + lifting) and the
# residual-only kernel
for nn in (1, 16, 2048):
    for beh in (jm.vonMisesIsotropicHardening(elasticity=el, yield_stress=jm.VoceHardening(sig0=350.0, sigu=500.0, b=1e3)),
                jm.GeneralIsotropicHardening(elasticity=el, yield_stress=jm.LinearHardening(sig0=200.0, H=10.0))):
        mh = jm.CUDAMaterial(beh); mh.set_data_manager(nn)
        mh.integrate(synth.strain(nn, 0, 1.25e-2, 1, 1)); mh.data_manager.update(); mh.integrate(synth.strain(nn, 0, 2e-2, 1, 1))
    mf = jm.CUDAMaterial(jm.FeFpJ2Plasticity(elasticity=el, yield_stress=jm.VoceHardening(sig0=500.0, sigu=750.0, b=1000.0)))
    mf.set_data_manager(nn); mf.integrate(synth.defgrad(nn, 0, 2e-2, 1, 1))
rowptr, colidx = nb.sparsity(ud, len(nodes))
bc, top = nb.boundary_conditions(nodes, 10.0)
system = AssembledSystem(ElementForms(ge, nb.W_DEG2), rowptr, colidx, bc=bc)
system.assemble(matrix=False); system.assemble(); system.get()
print("sanitize workload (last-session additions) ok")
