set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_default.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
python - <<'PY'
import csv, collections, re
rows = [r for r in csv.reader(open("gpurun_out/launches_bench_default.csv")) if len(r) > 5]
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[hi + 1:]:
    try: v = float(r[iv].replace(",", ""))
    except ValueError: continue
    v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(r[iu], 1e-6)
    name = re.sub(r"\(.*", "", r[ik]); tot[name] += v; cnt[name] += 1
s = sum(tot.values())
with open("gpurun_out/launches_bench_default_summary.txt", "w") as f:
    f.write("kernel launches of `python bench.py --steps 5 --warmup 3 --no-cpu` (ncu --metrics gpu__time_duration.sum, first 400 launches)\n\n")
    for k, v in tot.most_common():
        f.write(f"{k:62s} launches {cnt[k]:4d}  total {v:10.3f} ms  share {v / s:6.3f}\n")
print(open("gpurun_out/launches_bench_default_summary.txt").read())
PY
timeout 500 ncu --set full --clock-control none --import-source on -k regex:dxm_small_strain -s 10 -c 1 -o gpurun_out/j2_final -f python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_j2.log 2>&1; tail -2 gpurun_out/ncu_j2.log
python scripts/bench_latency.py > gpurun_out/latency.log 2>&1; tail -7 gpurun_out/latency.log
