#!/bin/bash
# round-end evidence: full GPU suite, smoke(), bench line, training-step profile, ncu launch list of the bench command
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1; echo "tests rc=$?"
tail -n 3 gpurun_out/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 4 gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['e2e']); print(d['roofline']['frac'], d.get('cpu_baseline'))
print({k:d['train'][k] for k in ('tiles_per_s','ms_per_step')}, d['wsi']['mpx_per_s'], d['bf16x3']['tiles_per_s'])
for k in d['kernels']:
    if 'conv3x3_tc' not in k['name']: print(k)
PY
timeout 300 python tools/train_profile.py > gpurun_out/train_profile.txt 2>&1; echo "train_profile rc=$?"
head -n 24 gpurun_out/train_profile.txt
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm rc=$?"; cat gpurun_out/bench_reference.json | cut -c1-400
bash tools/gpu_ncu_bench.sh 2>&1 | tail -n 14
