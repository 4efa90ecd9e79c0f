# One parameterised GPU session script (replaces the per-call scripts of round 1):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/gpu_call.sh <tag> <step> [<step> ...]'
# Every step writes gpurun_out/<tag>_<step>.*; copy what is cited into profiles/.  Steps:
#   tests      GPU parity suite              smoke     __graft_entry__.smoke()
#   bench      bench.py (default)            ref       bench.py --impl reference --steps 3 --warmup 1
#   fefp       scripts/bench_fefp.py 2e7     fefp_sus  sustained (>= 3 s) FeFp loop with clocks / power at 2e7
#   configs    scripts/bench_configs.py      launches  ncu launch list of the default bench
#   ncu_j2 | ncu_fefp | ncu_hosford | ncu_forms   one full ncu capture of that kernel family
#   latency    scripts/bench_latency.py      exchange  scripts/bench_exchange.py
#   forms      scripts/bench_fe_forms.py     any other word: run as `python scripts/<word>.py`
set -x
mkdir -p gpurun_out
tag=$1; shift
for step in "$@"; do
  o=gpurun_out/${tag}_${step}
  case $step in
    tests) python -m pytest tests -m gpu -x -q > $o.log 2>&1; echo "pytest rc=$?" >> $o.log; tail -4 $o.log ;;
    smoke) python -c "import __graft_entry__ as g; g.smoke()" > $o.log 2>&1; tail -5 $o.log ;;
    bench) python bench.py > $o.json 2> $o.err; cat $o.json ;;
    ref) python bench.py --impl reference --steps 3 --warmup 1 > $o.json 2> $o.err; cat $o.json ;;
    fefp) python scripts/bench_fefp.py 2e7 > $o.json 2>&1; cat $o.json ;;
    fefp_sus) python scripts/sweep_fefp_n.py 2e7 > $o.log 2>&1; cp gpurun_out/sweep_fefp_n.json $o.json; cat $o.log ;;
    configs) python scripts/bench_configs.py > $o.json 2> $o.err; tail -c 600 $o.json ;;
    launches) ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o.csv \
                python bench.py --steps 5 --warmup 3 --no-cpu > $o.log 2>&1 ;;
    ncu_j2) timeout 300 ncu --set full --clock-control none --import-source on -k regex:dxm_small_strain -s 10 -c 1 -o $o -f \
                python bench.py --steps 3 --warmup 3 --no-cpu > $o.log 2>&1; tail -2 $o.log ;;
    ncu_fefp) timeout 300 ncu --set full --clock-control none --import-source on -k regex:dxm_fefp -s 4 -c 1 -o $o -f \
                python scripts/bench_fefp.py 1e7 > $o.log 2>&1; tail -2 $o.log ;;
    ncu_hosford) timeout 300 ncu --set full --clock-control none --import-source on -k regex:dxm_hosford -s 2 -c 1 -o $o -f \
                python scripts/ncu_hosford.py > $o.log 2>&1; tail -3 $o.log ;;
    ncu_forms) timeout 300 ncu --set full --clock-control none --import-source on -k regex:fe_forms -s 1 -c 1 -o $o -f \
                python scripts/ncu_forms.py > $o.log 2>&1; tail -3 $o.log ;;
    latency) python scripts/bench_latency.py > $o.json 2> $o.err; cat $o.json ;;
    exchange) python scripts/bench_exchange.py > $o.json 2> $o.err; cat $o.json ;;
    forms) python scripts/bench_fe_forms.py > $o.json 2> $o.err; cat $o.json ;;
    *) python scripts/$step.py > $o.log 2>&1; tail -20 $o.log ;;
  esac
done
