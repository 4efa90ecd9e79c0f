set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -16 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -5 gpurun_out/smoke.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json
python scripts/bench_configs.py > gpurun_out/configs.log 2>&1
python - <<'PY'
import json
for r in json.load(open("gpurun_out/configs.json")):
    if "Hosford" in r["cfg"]:
        print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k not in ("note",)})
PY
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_workload.py > gpurun_out/sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/sanitizer_memcheck.log
