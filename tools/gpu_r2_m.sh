#!/bin/bash
# round 2, call M: full GPU suite, smoke(), full default bench (record), layer + train profiles
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
print({k:d['train'][k] for k in ('tiles_per_s','ms_per_step','wall_ms_per_step')}, d['bf16x3']['tiles_per_s'])
for k,w in d['wsi'].items(): print(k, w['mpx_per_s'], w['seconds_reps'], w['counts_tp_fp_fn_tn'])
PY
