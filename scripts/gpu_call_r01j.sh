set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
python scripts/bench_exchange.py 1e6 subset > gpurun_out/exchange_subset.log 2>&1; tail -30 gpurun_out/exchange_subset.log
python scripts/bench_configs.py > gpurun_out/configs.log 2>&1
python - <<'PY'
import json
for r in json.load(open("gpurun_out/configs.json")):
    if "divergence" in r["cfg"] or "faithful" in r["cfg"]:
        print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k not in ("note",)})
PY
tail -5 gpurun_out/configs.log
