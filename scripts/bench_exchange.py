"""cost of one QuadratureMap.update() around the CUDA material, wall clock, one B200.
(a) the reference's call sequence replayed verbatim (numpy gather, integrate -> pinned arrays, 3 NaN scans,
    fancy-index scatters -- tests/qmap_replay.py), (b) QuadratureExchange (contiguous fast path, direct DMA into
    the Function arrays, fused fail check, isv left on the GPU).  cfg5-like size: 1e6 P2 tets x 4 points."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dolfinx_materials_b200 as jm
from dolfinx_materials_b200.exchange import QuadratureExchange
from qmap_replay import QuadratureMapReplay
from oracle import synth

ncell, nqp = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000, 4
ntot = ncell * nqp
# "subset": the map covers 70 % of the cells (a material region of a multi-material mesh, multimaterials.py:265-273):
# the reference gathers / scatters with numpy fancy indexing, the exchange on the library's host thread pool
SUBSET = len(sys.argv) > 2 and sys.argv[2] == "subset"
cells = np.sort(np.random.default_rng(0).choice(ncell, int(0.7 * ncell), replace=False)) if SUBSET else None
nsub = (len(cells) if SUBSET else ncell) * nqp
out = []
for fefp in (False, True):
    el = jm.LinearElasticIsotropic(E=70e3, nu=0.3)
    mk = (lambda: jm.CUDAMaterial(jm.FeFpJ2Plasticity(elasticity=el, yield_stress=jm.VoceHardening(sig0=500.0, sigu=750.0, b=1000.0)))) if fefp else \
         (lambda: jm.CUDAMaterial(jm.vonMisesIsotropicHardening(elasticity=el, yield_stress=jm.VoceHardening(sig0=350.0, sigu=500.0, b=1e3))))
    gname, gdim = ("F", 9) if fefp else ("strain", 6)
    g0 = np.tile([1, 1, 1, 0, 0, 0, 0, 0, 0.0], (ntot, 1)) if fefp else np.zeros((ntot, 6))
    g1 = synth.defgrad(ntot, 1, 3e-2, 1, 1) if fefp else synth.strain(ntot, 1, 1.25e-2, 1, 1)
    # (a) reference sequence
    ref = QuadratureMapReplay(ncell, nqp, mk(), cells=cells); ref.register_gradient(gname, g0)
    if fefp: ref.update_initial_state("be_bar", np.array([1, 1, 1, 0, 0, 0.0]))
    ref.update(); ref.set_gradient_values(gname, g1); ref.update()
    t = []
    for _ in range(4):
        t0 = time.perf_counter(); ref.update(); t.append(time.perf_counter() - t0)
    ta = sorted(t)[1]
    # (b) exchange
    mat = mk(); grad = g0.copy().ravel(); flux = np.zeros(ntot * gdim)
    isv = {k: np.zeros(ntot * d) for k, d in mat.internal_state_variables.items()}; jac = np.zeros(ntot * gdim * gdim)
    ex = QuadratureExchange(mat, ncell, nqp, {gname: grad}, {mat.flux_names[0]: flux}, isv, jac, cells=cells)
    if fefp: ex.update_initial_state("be_bar", np.array([1, 1, 1, 0, 0, 0.0]))
    ex.update(); grad[:] = g1.ravel(); ex.update()
    t = []
    for _ in range(6):
        t0 = time.perf_counter(); s = ex.update(); t.append(time.perf_counter() - t0)
    tb = sorted(t)[2]
    assert np.array_equal(flux, ref.fluxes[mat.flux_names[0]].array) and np.array_equal(jac, ref.jacobian_flatten.array)
    t0 = time.perf_counter(); ex.advance(); tadv = time.perf_counter() - t0
    d2h = nsub * (gdim + gdim * gdim) * 8
    out.append(dict(behaviour="fefp" if fefp else "j2_voce", points=nsub, subset=SUBSET, reference_sequence_ms=ta * 1e3, exchange_ms=tb * 1e3,
                    speedup=ta / tb, exchange_gps=nsub / tb, exchange_d2h_gbs=d2h / tb / 1e9, kernel_ms=s.kernel_ms, advance_ms=tadv * 1e3))
    ex.close()
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True); json.dump(out, open("gpurun_out/exchange_subset.json" if SUBSET else "gpurun_out/exchange.json", "w"), indent=1)
