#!/bin/bash
# round 2, call S: per-variant resident weights (bres = 2: up1_conv1, up2_conv1): tests, per-layer profile on/off, bench A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_train.py -q -m gpu -x > gpurun_out/tests_s.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/tests_s.log
for f in 1 0; do
  ADP_RESIDENT_WEIGHTS=$f timeout 300 python tools/layer_profile.py > gpurun_out/layers_s_bres$f.txt 2>&1; head -n 1 gpurun_out/layers_s_bres$f.txt; grep -E "up1_conv1|up2_conv1" gpurun_out/layers_s_bres$f.txt
done
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --wsi none --no-x3"
for rep in 1 2; do
for f in 0 1; do
  ADP_RESIDENT_WEIGHTS=$f $B > gpurun_out/ab_s_${f}_$rep.json 2> gpurun_out/ab.err || tail -n 3 gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open('gpurun_out/ab_s_${f}_$rep.json'))
print('resident_weights=$f rep $rep train', round(d['train']['ms_per_step'],3), 'ms', round(d['train']['tiles_per_s'],1), 'tiles/s | infer', round(d['value'],2), d['clocks'].get('sm_mhz'), d['counts_tp_fp_fn_tn'])
PY
done
done
