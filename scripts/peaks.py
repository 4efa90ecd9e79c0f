"""Library microbenchmarks on one B200: copy, FP64 FMA, and the stream-mix ceilings."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dolfinx_materials_b200 import _lib
lib = _lib.load(); v = ctypes.c_double(); out = {}
_lib.check(lib.dxm_copy_peak(0, 1 << 32, ctypes.byref(v))); out["copy_gbs"] = v.value
_lib.check(lib.dxm_fp64_peak(0, ctypes.byref(v))); out["fp64_fma_tflops"] = v.value
for nr, nw, n in [(1, 1, 1 << 29), (37, 37, 40_000_000), (25, 49, 40_000_000), (25, 97, 20_000_000)]:
    _lib.check(lib.dxm_stream_peak(0, n, nr, nw, ctypes.byref(v))); out[f"stream_{nr}r_{nw}w_gbs"] = v.value
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True); json.dump(out, open("gpurun_out/peaks.json", "w"), indent=1)
