#!/bin/bash
# round 2, call A: build check, full GPU suite, smoke, short bench with the small WSI case
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 1500 python -m pytest tests -x -q -m gpu -s > gpurun_out/tests.log 2>&1; echo "tests rc=$?"
grep -E "configs\[1\]|1024\^2|ohem|passed|failed|Error|error" gpurun_out/tests.log | tail -n 40
tail -n 5 gpurun_out/tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 4 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 --wsi small --wsi-reps 2 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -n 5 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['e2e']); print(d['roofline']['frac'], d.get('cpu_baseline'))
print({k:d['train'][k] for k in ('tiles_per_s','ms_per_step','wall_ms_per_step')}); print(d['wsi'])
PY
