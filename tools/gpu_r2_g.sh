#!/bin/bash
# round 2, call G: full GPU suite after head epilogue / deterministic wgrad / CLI log changes; layer + train profile; short bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 2000 python -m pytest tests -q -m gpu -s > gpurun_out/tests.log 2>&1; echo "tests rc=$?"
grep -E "run-to-run|passed|failed|FAILED|Error" gpurun_out/tests.log | tail -n 20
timeout 300 python tools/layer_profile.py 1024 16 bf16 > gpurun_out/layers.txt 2>&1; grep -E "total|up1_conv3|down1_conv2|up1_conv2" gpurun_out/layers.txt
timeout 300 python tools/train_profile.py > gpurun_out/train_profile.txt 2>&1; head -n 12 gpurun_out/train_profile.txt
timeout 900 python bench.py --steps 10 --warmup 3 --wsi none --no-x3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -n 5 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['roofline']['frac'])
print({k:d['train'][k] for k in ('tiles_per_s','ms_per_step','wall_ms_per_step')})
PY
