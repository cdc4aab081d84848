#!/bin/bash
# ncu --set full of one training step's tensor-core kernels (20 forward convs, 20 data-gradient twins, 20 weight-gradient GEMMs)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
CMD="python tools/train_profile.py 1024 8 bf16"
$CMD > gpurun_out/plain_train.log 2>&1 && \
ncu --set full --clock-control none -k regex:"wgrad_tc_kernel|conv_tc_kernel" -s 300 -c 60 -o gpurun_out/prof_train $CMD > gpurun_out/ncu_train.log 2>&1
echo "train capture rc=$?"
ncu -i gpurun_out/prof_train.ncu-rep --page raw --csv > gpurun_out/prof_train_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/prof_train_raw.csv")))
hdr, units = rows[0], rows[1]
cols = ["ID", "Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size"]
idx = [hdr.index(c) for c in cols]
out = ["# ncu --set full --clock-control none -k regex:wgrad_tc_kernel|conv_tc_kernel -s 300 -c 60; cmd: python tools/train_profile.py 1024 8 bf16 (one training step, batch 8 at 1024^2)",
       ",".join(cols), ",".join(units[i] for i in idx)]
for r in rows[2:]:
    out.append(",".join('"' + r[i] + '"' if "," in r[i] else r[i] for i in idx))
open("gpurun_out/train_ncu_summary.csv", "w").write("\n".join(out) + "\n")
print("\n".join(out))
PY
rm -f gpurun_out/prof_train.ncu-rep gpurun_out/prof_train_raw.csv gpurun_out/prof_conv.ncu-rep gpurun_out/prof_hbm.ncu-rep
