set -x
mkdir -p gpurun_out
python -m pytest tests/test_hosford_gpu.py tests/test_full_size_gpu.py tests/test_exchange_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python scripts/bench_configs.py > gpurun_out/configs.log 2>&1
python - <<'PY'
import json
rows=[r for r in json.load(open("gpurun_out/configs.json")) if "Hosford a=10 alone" in r["cfg"]]
amps=sorted({r["cfg"].split("amp ")[1] for r in rows}, key=float)
print("split/direct | " + " | ".join(f"amp {a}" for a in amps))
for key in sorted({(r["split"], r["direct"]) for r in rows}):
    print(key, " | ".join(f'{[r["ms"] for r in rows if (r["split"], r["direct"])==key and r["cfg"].endswith("amp "+a)][0]:.3f}' for a in amps))
for r in json.load(open("gpurun_out/configs.json")):
    if "divergence" in r["cfg"] or "faithful" in r["cfg"]:
        print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()})
PY
tail -3 gpurun_out/configs.log
python scripts/bench_exchange.py 1e6 subset > gpurun_out/exchange_subset.log 2>&1; grep -E "exchange_ms|advance_ms|speedup" gpurun_out/exchange_subset.log
