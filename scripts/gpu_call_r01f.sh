set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python scripts/ab_host_mirror.py default 4 16 > gpurun_out/ab_host_mirror.log 2>&1
tail -20 gpurun_out/ab_host_mirror.log
python scripts/bench_configs.py > gpurun_out/configs.log 2>&1
tail -60 gpurun_out/configs.log
python scripts/ncu_hosford.py > gpurun_out/hosford_run.log 2>&1; cat gpurun_out/hosford_run.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dxm_hosford -s 2 -c 1 -o gpurun_out/hosford_full -f python scripts/ncu_hosford.py > gpurun_out/ncu_hosford.log 2>&1
tail -3 gpurun_out/ncu_hosford.log
DXM_HOST_MIRROR=0 python scripts/bench_exchange.py > gpurun_out/exchange_m0.log 2>&1; cp gpurun_out/exchange.json gpurun_out/exchange_m0.json
DXM_HOST_MIRROR=1 python scripts/bench_exchange.py > gpurun_out/exchange_m1.log 2>&1; cp gpurun_out/exchange.json gpurun_out/exchange_m1.json
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json
DXM_HOST_MIRROR=1 python bench.py --no-cpu --steps 10 > gpurun_out/bench_mirror.json 2> gpurun_out/bench_mirror.err; cat gpurun_out/bench_mirror.json
