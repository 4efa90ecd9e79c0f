set -x
mkdir -p gpurun_out
python -m pytest tests/test_hosford_gpu.py tests/test_full_size_gpu.py tests/test_qmap_replay_gpu.py tests/test_fmad_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python scripts/bench_configs.py > gpurun_out/configs.log 2>&1
python - <<'PY'
import json
for r in json.load(open("gpurun_out/configs.json")):
    if "Hosford" in r["cfg"]:
        print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k in ("cfg","split","ms","gps","plastic","hosford_ms","sorted_ms","shuffled_ms")})
PY
