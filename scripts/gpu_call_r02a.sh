# First GPU call of a new round (one B200, ~10 min of box time):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/gpu_call_r02a.sh'
# Re-establishes the baseline before anything is changed: GPU parity suite, smoke(), both bench arms, the launch list
# of the default bench and one full ncu capture each of the J2, FeFp and Hosford kernels (read them with
# `ncu -i gpurun_out/<name>.ncu-rep --page raw --csv`, copy what is cited into profiles/).
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -5 gpurun_out/smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json
python scripts/bench_fefp.py 2e7 > gpurun_out/fefp.json 2>&1; cat gpurun_out/fefp.json
python scripts/bench_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err; tail -c 1500 gpurun_out/configs.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_default.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dxm_small_strain -s 10 -c 1 -o gpurun_out/j2 -f \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_j2.log 2>&1; tail -2 gpurun_out/ncu_j2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dxm_fefp -s 4 -c 1 -o gpurun_out/fefp -f \
    python scripts/bench_fefp.py 1e7 > gpurun_out/ncu_fefp.log 2>&1; tail -2 gpurun_out/ncu_fefp.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dxm_hosford -s 2 -c 1 -o gpurun_out/hosford -f \
    python scripts/ncu_hosford.py > gpurun_out/ncu_hosford.log 2>&1; tail -3 gpurun_out/ncu_hosford.log
