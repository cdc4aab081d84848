#!/bin/bash
# ncu evidence: (1) launch list with per-launch device time, (2) full capture of the conv kernel instances
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
# (1) launch list of the bench command itself (all launches of warm-up, timed, e2e and profile steps)
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --wsi none --train-batch 0"
$BENCH > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_bench.csv $BENCH > gpurun_out/ncu0.log 2>&1
echo "bench launch list rc=$?"
CMD="python tools/layer_profile.py 1024 8 bf16"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 40 -c 20 -o gpurun_out/prof_conv $CMD > gpurun_out/ncu2.log 2>&1
echo "full capture rc=$?"
tail -n 3 gpurun_out/ncu1.log gpurun_out/ncu2.log
ls -la gpurun_out/
