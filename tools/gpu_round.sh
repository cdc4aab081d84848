#!/bin/bash
# one box: build, full GPU suite, training-step profile, bench (+ TTA A/B via ADP_TTA_SERIAL)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1; echo "tests rc=$?"
tail -n 4 gpurun_out/tests.log
timeout 300 python tools/train_profile.py > gpurun_out/train_profile.txt 2>&1; echo "train_profile rc=$?"
head -n 24 gpurun_out/train_profile.txt
timeout 600 python bench.py --steps 5 --warmup 3 $BENCH_ARGS > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['e2e']); print(d['roofline']['frac'], d.get('cpu_baseline'))
print({k:d['train'][k] for k in ('tiles_per_s','ms_per_step')}, {k: w['mpx_per_s'] for k, w in d['wsi'].items()})
for k in d['kernels']:
    if 'conv3x3_tc' not in k['name']: print(k)
PY
if [ "$TTA_AB" == "1" ]; then
  ADP_TTA_SERIAL=1 timeout 300 python tools/layer_profile.py 1024 16 bf16 tta 2>&1 | grep -E "tta|total"
  timeout 300 python tools/layer_profile.py 1024 16 bf16 tta 2>&1 | grep -E "tta|total|maxpool|add6|first_conv"
fi
