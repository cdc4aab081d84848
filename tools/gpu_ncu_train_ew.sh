#!/bin/bash
# ncu --set full of one training step's elementwise / CUDA-core kernels (max-pool backward, upsample materialise / backward,
# dropout, first-layer weight gradient, head backward, loss reduce)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
CMD="python tools/train_profile.py 1024 8 bf16"
# KREGEX / SKIP / COUNT select other kernels (defaults: 17 launches of the sixth step = the profiled one)
KREGEX=${KREGEX:-"maxpool2_bwd_kernel|upsample2_kernel|upsample2_bwd_kernel|dropout_kernel|dropout_dense_kernel|first_wgrad_kernel|head_bwd_kernel|loss_reduce_kernel|first_conv_kernel"}
ncu --set full --clock-control none -k regex:"$KREGEX" -s ${SKIP:-85} -c ${COUNT:-17} -o gpurun_out/prof_ew $CMD > gpurun_out/ncu_ew.log 2>&1
echo "capture rc=$?"
ncu -i gpurun_out/prof_ew.ncu-rep --page raw --csv > gpurun_out/prof_ew_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/prof_ew_raw.csv")))
hdr, units = rows[0], rows[1]
cols = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]
idx = [hdr.index(c) for c in cols if c in hdr]
stall = [i for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
out = ["# ncu --set full --clock-control none; cmd: python tools/train_profile.py 1024 8 bf16 (one training step, batch 8 at 1024^2)",
       ",".join(hdr[i] for i in idx) + ",top stalls (warps per issue)", ",".join(units[i] for i in idx) + ","]
for r in rows[2:]:
    st = sorted(((float(r[i].replace(",", "")) if r[i] else 0.0, hdr[i][len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for i in stall), reverse=True)[:3]
    out.append(",".join('"' + r[i] + '"' if "," in r[i] else r[i] for i in idx) + "," + " ".join(f"{n}={v:.1f}" for v, n in st))
open("gpurun_out/train_ew_ncu_summary.csv", "w").write("\n".join(out) + "\n")
print("\n".join(out))
PY
rm -f gpurun_out/prof_ew.ncu-rep gpurun_out/prof_ew_raw.csv
