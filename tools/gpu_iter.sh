#!/bin/bash
# quick iteration: bf16 parity tests + per-layer profile (+ optional bench)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_forward.py -x -q -m gpu -k "bf16" > gpurun_out/tests_bf16.log 2>&1; echo "tests_bf16 rc=$?"
tail -n 4 gpurun_out/tests_bf16.log
timeout 300 python tools/layer_profile.py 1024 16 bf16 2>&1 | tee gpurun_out/layers.txt
if [ "$1" == "bench" ]; then timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['e2e'], d['roofline'])
PY
fi
