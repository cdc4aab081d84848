#!/bin/bash
# round 2, call C: full GPU suite, then the default bench (full WSI cases)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 2000 python -m pytest tests -q -m gpu -s > gpurun_out/tests.log 2>&1; echo "tests rc=$?"
grep -E "trained|passed|failed|FAILED|Error" gpurun_out/tests.log | tail -n 30
( time timeout 1500 python bench.py --steps 10 --warmup 3 $BENCH_ARGS > gpurun_out/bench.json 2> gpurun_out/bench.err ) 2>&1 | tail -n 4; tail -n 5 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['e2e']); print(d['roofline']['frac'], d.get('cpu_baseline'))
print({k:d['train'][k] for k in ('tiles_per_s','ms_per_step','wall_ms_per_step')}); print(json.dumps(d['wsi'], indent=1))
PY
