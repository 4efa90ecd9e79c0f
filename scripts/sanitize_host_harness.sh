# The kernels' per-point / per-cell routines (host-compiled, tests/*_host_check.cu) under AddressSanitizer + UBSan on
# the CPU: builds instrumented copies of the three harness libraries, runs the host-harness tests against them, restores
# the plain builds.  Expected: all tests pass, no "runtime error" / "AddressSanitizer" line.   bash scripts/sanitize_host_harness.sh
set -e
cd "$(dirname "$0")/.."
out=$(mktemp -d)
mkdir -p tests/_build "$out/keep"
python -m pytest tests/test_point_host.py tests/test_hosford_host.py tests/test_fe_host.py -q -x > /dev/null  # plain builds exist
cp tests/_build/lib*_host_check.so "$out/keep/"
for f in point_host_check hosford_host_check fe_host_check; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 -fmad=false \
       -Xcompiler -fPIC,-ffp-contract=off,-fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer \
       -shared -o tests/_build/lib$f.so tests/$f.cu
done
asan=$(gcc -print-file-name=libasan.so); ubsan=$(gcc -print-file-name=libubsan.so)
set +e
LD_PRELOAD="$asan $ubsan" ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 \
  python -m pytest tests/test_point_host.py tests/test_hosford_host.py tests/test_fe_host.py -x -q -s -p no:cacheprovider > "$out/run.log" 2>&1
rc=$?
set -e
cp "$out/keep/"*.so tests/_build/; touch tests/_build/*.so
echo "pytest rc=$rc; UBSan reports: $(grep -c 'runtime error' "$out/run.log" || true); ASan reports: $(grep -c 'AddressSanitizer' "$out/run.log" || true)"
tail -1 "$out/run.log"
