#!/usr/bin/env python
"""Summarise an ncu --set full report (run here, no GPU needed): one row per launch with the metrics the roofline discussion
uses, plus the per-forward DRAM traffic json that bench.py's roofline.traffic reads.

  python tools/ncu_summarise.py gpurun_out/prof_conv.ncu-rep profiles/r2_ncu_conv.csv [profiles/r2_ncu_conv_traffic.json FORWARDS_PER_LAUNCH] [layer names...]
"""
import csv
import io
import json
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_uniform.sum.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active"]

LAYERS = ["down1_conv2", "down2_conv1", "down2_conv2", "down3_conv1", "down3_conv2", "dilate1", "dilate2", "dilate3", "dilate4", "dilate5",
          "dilate6", "up3_conv1", "up3_conv2", "up3_conv3", "up2_conv1", "up2_conv2", "up2_conv3", "up1_conv1", "up1_conv2", "up1_conv3"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {m: hdr.index(m) for m in METRICS if m in hdr}
    ik = hdr.index("Kernel Name")
    names = sys.argv[5:] if len(sys.argv) > 5 else (LAYERS if len(data) == len(LAYERS) else None)
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["# " + " ".join(sys.argv)])
        w.writerow(["layer", "ID", "Kernel Name"] + list(idx))
        w.writerow(["", "", ""] + [units[i] for i in idx.values()])
        for n, r in enumerate(data):
            w.writerow([names[n] if names and n < len(names) else "", n, r[ik]] + [r[i] for i in idx.values()])
    print("wrote", out, len(data), "launches")
    if len(sys.argv) > 4:
        fw = int(sys.argv[4])
        def tobytes(v, u):
            v = float(v.replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        rd = sum(tobytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) for r in data) / fw
        wr = sum(tobytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]]) for r in data) / fw
        json.dump({"source": f"{out} (ncu --set full, {len(data)} conv_tc_kernel launches = one chunk of {fw} forwards at 1024^2, bf16)",
                   "forwards_per_launch_in_capture": fw, "launches": len(data), "dram_bytes_read_per_forward": rd,
                   "dram_bytes_write_per_forward": wr, "dram_bytes_per_forward": rd + wr}, open(sys.argv[3], "w"), indent=1)
        print("wrote", sys.argv[3], (rd + wr) / 1e9, "GB per forward")


if __name__ == "__main__":
    main()
