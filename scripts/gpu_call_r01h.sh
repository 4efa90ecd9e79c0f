set -x
mkdir -p gpurun_out
python -m pytest tests/test_hosford_gpu.py tests/test_fmad_gpu.py tests/test_qmap_replay_gpu.py tests/test_host_mirror_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python scripts/bench_configs.py > gpurun_out/configs.log 2>&1
python - <<'PY'
import json
for r in json.load(open("gpurun_out/configs.json")):
    if "Hosford" in r["cfg"]:
        print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k not in ("note",)})
PY
DXM_FMAD=1 python scripts/fmad_check.py 2>/dev/null | tail -1 > gpurun_out/fmad_check.json; cat gpurun_out/fmad_check.json
DXM_HOS_SPLIT=2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:dxm_hosford -s 2 -c 1 -o gpurun_out/hosford_v4_tiled -f python scripts/ncu_hosford.py > gpurun_out/ncu_hosford_tiled.log 2>&1
tail -2 gpurun_out/ncu_hosford_tiled.log
