#!/bin/bash
# ncu --set full of the HBM-bound kernels of the path (first conv, max-pool, Add, TTA combine / blend, finalize): achieved DRAM
# bytes and throughput per launch, from the bench command (1024^2 tiles, 16 forwards per launch) with a small WSI run appended
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --wsi small --train-batch 0"
$CMD > gpurun_out/plain_hbm.log 2>&1 && \
ncu --set full --clock-control none -k regex:"first_conv_kernel|add6_kernel|maxpool2_kernel|tta_blend_kernel|finalize_kernel" -s 12 -c 40 \
    -o gpurun_out/prof_hbm $CMD > gpurun_out/ncu_hbm.log 2>&1
echo "hbm capture rc=$?"
ncu -i gpurun_out/prof_hbm.ncu-rep --page raw --csv > gpurun_out/prof_hbm_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/prof_hbm_raw.csv")))
hdr, units = rows[0], rows[1]
cols = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size"]
idx = [hdr.index(c) for c in cols if c in hdr]
print(",".join(hdr[i] for i in idx)); print(",".join(units[i] for i in idx))
for r in rows[2:]:
    print(",".join('"' + r[i] + '"' if "," in r[i] else r[i] for i in idx))
PY
