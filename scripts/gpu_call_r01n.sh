set -x
mkdir -p gpurun_out
python -m pytest tests/test_exchange_gpu.py tests/test_qmap_replay_gpu.py tests/test_api_corners_gpu.py tests/test_small_strain_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python scripts/bench_exchange.py 1e6 subset > gpurun_out/exchange_subset.log 2>&1; grep -E "behaviour|exchange_ms|advance_ms|speedup|reference_sequence" gpurun_out/exchange_subset.log
python scripts/bench_exchange.py 1e6 > gpurun_out/exchange_full.log 2>&1; grep -E "behaviour|exchange_ms|advance_ms|speedup" gpurun_out/exchange_full.log
