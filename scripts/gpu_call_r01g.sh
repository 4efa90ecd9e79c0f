set -x
mkdir -p gpurun_out
python -m pytest tests/test_hosford_gpu.py tests/test_small_strain_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python scripts/bench_configs.py > gpurun_out/configs.log 2>&1
python - <<'PY'
import json
for r in json.load(open("gpurun_out/configs.json")):
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k not in ("note",)})
PY
python scripts/ncu_hosford.py > gpurun_out/hosford_run.log 2>&1; cat gpurun_out/hosford_run.log
DXM_HOS_SPLIT=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:dxm_hosford -s 4 -c 2 -o gpurun_out/hosford_v3_full -f python scripts/ncu_hosford.py > gpurun_out/ncu_hosford.log 2>&1
DXM_HOS_SPLIT=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:dxm_hosford -s 2 -c 1 -o gpurun_out/hosford_v3_fused -f python scripts/ncu_hosford.py > gpurun_out/ncu_hosford_fused.log 2>&1
tail -3 gpurun_out/ncu_hosford.log
timeout 600 compute-sanitizer --tool racecheck python scripts/sanitize_workload.py > gpurun_out/sanitizer_racecheck.log 2>&1; tail -4 gpurun_out/sanitizer_racecheck.log
