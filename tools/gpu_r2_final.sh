#!/bin/bash
# round 2 final evidence: full GPU suite, smoke(), profiles, full default bench, reference arm, ncu launch list + full capture of the conv launches
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 2000 python -m pytest tests -q -m gpu > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
timeout 300 python tools/layer_profile.py 1024 16 bf16 > gpurun_out/layers.txt 2>&1; grep -E "total" gpurun_out/layers.txt
timeout 300 python tools/train_profile.py > gpurun_out/train_profile.txt 2>&1; head -n 2 gpurun_out/train_profile.txt
( time timeout 1500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err ) 2>&1 | tail -n 4; tail -n 3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['e2e']['value'], d['roofline']['frac'], d.get('cpu_baseline',{}).get('value'))
print({k:d['train'][k] for k in ('tiles_per_s','ms_per_step','wall_ms_per_step','tflops_3x_forward')}, d['bf16x3']['tiles_per_s'], d['first_conv_fusion']['tiles_per_s'])
for k,w in d['wsi'].items(): print(k, w['mpx_per_s'], w['seconds_reps'], w['counts_tp_fp_fn_tn'])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm rc=$?"; cut -c1-300 gpurun_out/bench_reference.json
BENCH="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --wsi none --train-batch 0 --no-x3"
$BENCH > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_bench.csv $BENCH > gpurun_out/ncu0.log 2>&1
echo "bench launch list rc=$?"
CMD="python tools/layer_profile.py 1024 8 bf16"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 40 -c 20 -f -o gpurun_out/prof_conv_r2 $CMD > gpurun_out/ncu2.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/ | tail -n 8
