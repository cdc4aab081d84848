#!/bin/bash
# quick iteration: parity tests + per-layer profile (+ optional bench)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
KSEL=${KSEL:-bf16}
timeout 900 python -m pytest tests/test_gpu_forward.py -x -q -m gpu -k "$KSEL" > gpurun_out/tests_fwd.log 2>&1; echo "tests_fwd rc=$?"
tail -n 6 gpurun_out/tests_fwd.log
timeout 300 python tools/layer_profile.py 1024 16 bf16 2>&1 | tee gpurun_out/layers.txt
if [ "$1" == "bench" ]; then timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['e2e'], d['roofline'])
PY
fi
