"""flux/tangent (device) -> element forms / assembled CSR system, one B200.
P2 tets x 4 quadrature points (the cfg5 element), FeFp behaviour.  Reports the fused device assembly against the
transfer it replaces (the (n, 81) tangent + (n, 9) flux D2H that feeds DOLFINx's assembler in the reference)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dolfinx_materials_b200 as jm
from dolfinx_materials_b200.fe import AssembledSystem, ElementForms, GradientEvaluator
from dolfinx_materials_b200.material import PinnedArray
from oracle import fe_forms as ff
from oracle import fe_gradient as fg

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 48
order = int(sys.argv[2]) if len(sys.argv) > 2 else 2
coords, gd, ud, nodes = fg.box_tets(nx, nx, nx, order)
qp = fg.TET_QP_DEG2 if order == 2 else fg.TET_QP_DEG1
dphi = fg.tet_dphi(qp, order)
w = np.full(len(qp), 1.0 / 6.0 / len(qp))
nc, nqp = len(gd), len(qp)
n = nc * nqp
mat = jm.CUDAMaterial(jm.FeFpJ2Plasticity(elasticity=jm.LinearElasticIsotropic(E=70e3, nu=0.3),
                                          yield_stress=jm.VoceHardening(sig0=500.0, sigu=750.0, b=1000.0)))
mat.set_data_manager(n)
ge = GradientEvaluator(mat, coords, gd, ud, dphi)
forms = ElementForms(ge, w)
t0 = time.perf_counter(); rowptr, colidx = ff.sparsity(ud, len(nodes), 3); t_pat = time.perf_counter() - t0
system = AssembledSystem(forms, rowptr, colidx, bc=np.repeat(nodes[:, 0] < 0.04, 3))
x, y, z = nodes.T
u = (0.02 * np.stack([x * y + 0.5 * z * z, -2 * y * z + 0.3 * x * x, 0.7 * x * z - 0.4 * y * y], axis=1)).ravel()
ge.eval(u)
s = mat.integrate_resident()


def timeit(fn, reps=10):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return sorted(ts)[reps // 3]


t_asm = timeit(lambda: system.assemble())
t_vec = timeit(lambda: system.assemble(matrix=False))
if os.environ.get("DXM_FORMS_QUICK"):  # kernel A/B runs (scripts/ab_variants.py): the two assembly times only
    out = dict(cells=nc, assemble_matrix_and_vector_ms=t_asm * 1e3, assemble_vector_only_ms=t_vec * 1e3)
    print(json.dumps(out)); os.makedirs("gpurun_out", exist_ok=True); json.dump(out, open("gpurun_out/fe_forms.json", "w"))
    sys.exit(0)
t_get = timeit(lambda: system.get(), 3)
ndof = forms.ndof
ke_pin = PinnedArray((nc, ndof, ndof)); fe_pin = PinnedArray((nc, ndof))
t_elem = timeit(lambda: forms.compute(out_fe=fe_pin.array, out_ke=ke_pin.array), 3)
# what the reference sequence moves instead: flux + tangent D2H through the host-facing integrate
ct_pin = PinnedArray((n, 81)); fl_pin = PinnedArray((n, 9))
t_d2h = timeit(lambda: (mat.read_state_into("Ct", ct_pin.array), mat.read_state_into("PK1", fl_pin.array)), 3)
# oracle (numpy) on a subset of cells, extrapolated
sub = min(nc, 4096)
flux = np.ascontiguousarray(mat.device_view("PK1").cpu().numpy().T)[: sub * nqp]
ct = np.ascontiguousarray(mat.device_tangent().cpu().numpy().T)[: sub * nqp]
t0 = time.perf_counter(); fe_ref, ke_ref = ff.element_forms(coords, gd[:sub], ud[:sub], dphi, w, flux, ct, 1, 3); t_or = time.perf_counter() - t0
assert np.array_equal(ke_pin.array[:sub], ke_ref) and np.array_equal(fe_pin.array[:sub], fe_ref)
ct_bytes = n * 81 * 8
out = dict(cells=nc, points=n, dofs=len(nodes) * 3, nnz=int(system.nnz), order=order, plastic=s.n_plastic / n,
           assemble_matrix_and_vector_ms=t_asm * 1e3, assemble_vector_only_ms=t_vec * 1e3,
           assemble_cells_per_s=nc / t_asm, tangent_read_gbs=ct_bytes / t_asm / 1e9,
           system_d2h_ms=t_get * 1e3, system_bytes=int(system.nnz) * 8 + len(nodes) * 24,
           element_forms_to_pinned_host_ms=t_elem * 1e3, element_bytes=nc * ndof * (ndof + 1) * 8,
           flux_tangent_d2h_ms_replaced=t_d2h * 1e3, flux_tangent_bytes=n * 90 * 8,
           numpy_oracle_cells_per_s=sub / t_or, pattern_build_s_scipy=t_pat)
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True); json.dump(out, open("gpurun_out/fe_forms.json", "w"), indent=1)
