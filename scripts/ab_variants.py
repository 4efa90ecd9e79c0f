"""Kernel experiments side by side on one box: runs a measurement script once per experiment build
(``lib/libdxm_cuda_<name>.so``, made with ``DXM_VARIANT=<name> DXM_VARIANT_DEFS=... python -m dolfinx_materials_b200.build``)
and once with the product library, each in its own process.

    python scripts/ab_variants.py [script=scripts/ab_hosford.py] [args...]

Every run's JSON (``gpurun_out/<script>.json``) is collected into ``gpurun_out/ab_variants.json`` under the variant's name
('' = product library)."""
import glob, json, os, subprocess, sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
script = sys.argv[1] if len(sys.argv) > 1 else "scripts/ab_hosford.py"
args = sys.argv[2:]
libs = sorted(glob.glob(os.path.join(root, "dolfinx_materials_b200", "lib", "libdxm_cuda_*.so")))
names = [""] + [os.path.basename(p)[len("libdxm_cuda_"):-3] for p in libs if not p.endswith("_unfused.so")]
res = {}
produced = os.path.join(root, "gpurun_out", os.path.basename(script)[:-3] + ".json")
for name in names:
    env = dict(os.environ, DXM_VARIANT=name)
    env.pop("DXM_UNFUSED", None)
    print(f"==== variant '{name}'", flush=True)
    t0 = __import__("time").time()
    r = subprocess.run([sys.executable, os.path.join(root, script), *args], env=env, cwd=root, stdout=subprocess.DEVNULL)
    if not os.path.exists(produced) or os.path.getmtime(produced) < t0:  # the script's own file name: newest JSON
        new = [p for p in glob.glob(os.path.join(root, "gpurun_out", "*.json")) if os.path.getmtime(p) >= t0]
        produced = max(new, key=os.path.getmtime) if new else produced
    if r.returncode == 0 and os.path.exists(produced):
        res[name or "product"] = json.load(open(produced))
    else:
        res[name or "product"] = {"error": r.returncode}
json.dump(res, open(os.path.join(root, "gpurun_out", "ab_variants.json"), "w"), indent=1)
