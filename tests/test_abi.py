"""The C-ABI library builds for sm_100a without a GPU, exports every symbol include/dxm.h declares,
fails loudly without a device, and the product package never touches the oracle (CPU only)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "dxm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dxm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(jm):
    from dolfinx_materials_b200 import _lib

    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/dxm.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes signatures and header drifted apart"
    assert b"sm_100a" in lib.dxm_version()


def test_integration_guide_names_every_entry_point():
    """INTEGRATION.md section 3 maps each C entry point to the reference interface it replaces: none may be missing."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [s for s in header_symbols() if s not in doc]
    assert not missing, f"INTEGRATION.md does not mention {missing}"


def test_library_contains_sm100a_sass_only(jm):
    from dolfinx_materials_b200 import _lib

    out = subprocess.run(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback(jm):
    """Without a CUDA device the product path raises (it never routes through the oracle)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = jm.CUDAMaterial(jm.ElasticBehavior(elasticity=jm.LinearElasticIsotropic(E=1.0, nu=0.2)))
    with pytest.raises(RuntimeError):
        m.set_data_manager(8)
    with pytest.raises(RuntimeError):
        m.integrate([[0.0] * 6] * 8)


def test_experiment_build_selection_fails_loudly_when_missing():
    """``DXM_VARIANT=<name>`` selects ``lib/libdxm_cuda_<name>.so`` (kernel A/B runs, scripts/ab_variants.py); a variant
    that was never built raises at the first use instead of silently loading the product library."""
    code = ("import dolfinx_materials_b200 as jm, sys\n"
            "from dolfinx_materials_b200 import _lib\n"
            "assert _lib.LIB_PATH.name == 'libdxm_cuda_doesnotexist.so', _lib.LIB_PATH\n"
            "try:\n    _lib.load()\nexcept RuntimeError as e:\n    assert 'missing' in str(e); sys.exit(0)\n"
            "sys.exit(1)\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, DXM_VARIANT="doesnotexist", PYTHONPATH=root)
    env.pop("DXM_UNFUSED", None)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=root)
    assert r.returncode == 0, r.stderr


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dolfinx_materials_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f


def test_protocol_surface_without_gpu(jm):
    """Names, sizes and ordering the reference's QuadratureMap relies on (quadrature_map.py:84-137,
    :338-348; jaxmat.py:166-193)."""
    el = jm.LinearElasticIsotropic(E=70e3, nu=0.3)
    m = jm.CUDAMaterial(jm.vonMisesIsotropicHardening(elasticity=el, yield_stress=jm.VoceHardening(sig0=350.0, sigu=500.0, b=1e3)))
    assert m.gradients == {"strain": 6} and m.fluxes == {"stress": 6}
    assert list(m.internal_state_variables.items()) == [("p", 1), ("epsp", 6)]
    assert m.tangent_blocks == {("stress", "strain"): (6, 6)}
    assert m.gradient_names == ["strain"] and m.flux_names == ["stress"]
    assert m.rotation_matrix is None and m.name == "vonMisesIsotropicHardening"
    assert set(m.material_properties) == {"E", "nu", "sig0", "sigu", "b", "H"}
    f = jm.CUDAMaterial(jm.FeFpJ2Plasticity(elasticity=el, yield_stress=jm.VoceHardening(sig0=500.0, sigu=750.0, b=1000.0)))
    assert f.gradients == {"F": 9} and f.fluxes == {"PK1": 9}
    assert list(f.internal_state_variables.items()) == [("p", 1), ("be_bar", 6)]
    assert f.tangent_blocks == {("PK1", "F"): (9, 9)}
    g = jm.CUDAMaterial(jm.GeneralIsotropicHardening(elasticity=el, yield_stress=jm.LinearHardening(sig0=200.0, H=10.0)))
    assert g.gradients == {"strain": 6} and g.fluxes == {"stress": 6} and g.tangent_blocks == {("stress", "strain"): (6, 6)}
    assert g.material_properties == {"E": 70e3, "nu": 0.3, "sig0": 200.0, "H": 10.0, "a": 10}  # the demo's exponent
    with pytest.raises(ValueError):
        jm.GeneralIsotropicHardening(elasticity=el, yield_stress=jm.LinearHardening(sig0=1.0), equivalent_stress=jm.Hosford(a=7))
    gv = jm.CUDAMaterial(jm.GeneralIsotropicHardening(elasticity=el, yield_stress=jm.VoceHardening(sig0=1.0, sigu=2.0, b=1.0)))
    assert set(gv.material_properties) == {"E", "nu", "sig0", "sigu", "b", "H", "a"}
    with pytest.raises(TypeError):
        jm.GeneralIsotropicHardening(elasticity=el, yield_stress=jm.TabulatedHardening(p=[0.0, 1.0], sig=[1.0, 2.0]))
    with pytest.raises(TypeError):
        jm.CUDAMaterial(jm.vonMisesIsotropicHardening(elasticity=el, yield_stress=lambda p: 1.0 + p ** 0.3))  # not of the kernels' family
    with pytest.raises(KeyError):
        m.update_material_property("nope", 1.0)


def test_packed_tangent_index_map():
    """SYM6_PACKED (host mirror of sym6_packed in csrc/dxm_canon.cuh): the 21 rows of the resident small-strain
    tangent, upper triangle row-major, shared by (j, i) and (i, j)."""
    from dolfinx_materials_b200.material import SYM6_PACKED

    m = SYM6_PACKED.reshape(6, 6)
    assert np.array_equal(m, m.T) and sorted(set(m.ravel())) == list(range(21))
    assert [m[j, i] for j in range(6) for i in range(j, 6)] == list(range(21))


@pytest.mark.parametrize("n,threads", [(0, 1), (1, 1), (7, 0), (5000, 1), (40001, 0), (40001, 3)])
def test_host_mirror_of_the_packed_tangent(jm, n, threads):
    """dxm_host_mirror_sym6 (host half of the packed-tangent hand-off, no GPU involved): (n, 21) -> (n, 36), every
    value copied bit for bit, on aligned (streaming stores) and unaligned destinations, 1 thread and the pool."""
    from dolfinx_materials_b200 import _lib
    from dolfinx_materials_b200.material import SYM6_PACKED

    lib = _lib.load()
    rng = np.random.default_rng(n)
    packed = rng.standard_normal((n, 21))
    if n:
        packed[0, :3] = [np.nan, -0.0, np.inf]
    want = packed[:, SYM6_PACKED]
    for offset in (0, 1):  # offset 1 double: 8-byte aligned only -> plain-store path
        buf = np.full(n * 36 + 2 + offset, -7.0)
        base = (-buf.ctypes.data // 8) % 2  # make `full` 16-byte aligned, then shift by `offset`
        full = buf[base + offset: base + offset + n * 36]
        assert (full.ctypes.data % 16 == 0) == (offset == 0) or n == 0
        rc = lib.dxm_host_mirror_sym6(packed.ctypes.data_as(ctypes.c_void_p), full.ctypes.data_as(ctypes.c_void_p), n, threads)
        assert rc == 0
        assert full.reshape(n, 36).tobytes() == want.tobytes()
        assert buf[base + offset + n * 36] == -7.0 and (base + offset == 0 or buf[base + offset - 1] == -7.0)


@pytest.mark.parametrize("n,row_len,threads", [(0, 5, 1), (1, 1, 1), (1000, 24, 1), (30000, 36, 0), (30000, 7, 3)])
def test_host_row_gather_and_scatter(jm, n, row_len, threads):
    """dxm_host_gather_rows / dxm_host_scatter_rows == numpy fancy indexing (the subset passes of QuadratureMap)."""
    from dolfinx_materials_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(n + row_len)
    total = 3 * n + 5
    rows = np.sort(rng.choice(total, size=n, replace=False)).astype(np.int64)
    src = rng.standard_normal((total, row_len))
    dst = np.full((n, row_len), np.nan)
    c = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    assert lib.dxm_host_gather_rows(c(src), c(rows), n, row_len, c(dst), threads) == 0
    assert np.array_equal(dst, src[rows])
    big = np.zeros((total, row_len))
    vals = rng.standard_normal((n, row_len))
    assert lib.dxm_host_scatter_rows(c(big), c(rows), n, row_len, c(vals), threads) == 0
    want = np.zeros((total, row_len))
    want[rows] = vals
    assert np.array_equal(big, want)
