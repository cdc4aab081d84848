#!/bin/bash
# round 2, call D: full GPU suite (continue past failures) after the I/O front-end / hygiene changes, short bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 2000 python -m pytest tests -q -m gpu -s > gpurun_out/tests.log 2>&1; echo "tests rc=$?"
grep -E "nvJPEG|TIFF-LZW|prediction from|recon nvJPEG|passed|failed|FAILED|Error" gpurun_out/tests.log | tail -n 40
timeout 900 python bench.py --steps 5 --warmup 3 --wsi none --no-x3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -n 5 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['roofline']['frac'])
print({k:d['train'][k] for k in ('tiles_per_s','ms_per_step','wall_ms_per_step')})
PY
