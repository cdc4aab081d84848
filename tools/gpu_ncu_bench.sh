#!/bin/bash
# ncu launch list of the bench command itself (per-launch gpu__time_duration, serialised, cold cache): the kernels' SHARES of a
# step are what bench.py's roofline.share_of_step must agree with
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --wsi none --train-batch 0"
$BENCH > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_bench.csv $BENCH > gpurun_out/ncu0.log 2>&1
echo "bench launch list rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_bench.csv")) if len(r) > 5]
hdr = rows[0]
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
agg = collections.OrderedDict()
n = 0
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    name = r[ik].split("(")[0].replace("void ", "").replace("adp::", "")
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[iv].replace(",", ""))
    n += 1
tot = sum(v[1] for v in agg.values())
print(f"{n} launches, {tot/1e6:.2f} ms total (ncu-serialised)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} n={v[0]:5d} {v[1]/1e6:9.3f} ms {100*v[1]/tot:6.2f}%")
PY
