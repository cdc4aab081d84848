#!/bin/bash
# round 2, call F: training tests + A/B of the side-stream overlap in the backward pass
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_dp.py -q -m gpu 2>&1 | tail -n 5
for ov in 1 0 1 0; do
  ADP_TRAIN_OVERLAP=$ov timeout 600 python bench.py --steps 20 --warmup 3 --wsi none --no-x3 --no-cpu-baseline > gpurun_out/bench_ov$ov.json 2> gpurun_out/bench.err || tail -n 5 gpurun_out/bench.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_ov$ov.json'))
print("overlap=$ov", {k:d['train'][k] for k in ('tiles_per_s','ms_per_step','wall_ms_per_step')}, 'fwd', d['value'], d['clocks'])
PY
done
