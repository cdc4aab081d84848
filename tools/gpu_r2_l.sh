#!/bin/bash
# round 2, call L: first-conv fusion with row carry: bit-identity, layer profile, bench A/B
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_forward.py -q -m gpu -x -k "fusion" -s > gpurun_out/tests_fc.log 2>&1; rc=$?; echo "fc test rc=$rc"; tail -n 5 gpurun_out/tests_fc.log
if [ $rc -ne 0 ]; then timeout 600 compute-sanitizer --tool memcheck python tools/fc_probe.py 256 > gpurun_out/sanitizer_fc.log 2>&1; grep -v "^=========     " gpurun_out/sanitizer_fc.log | head -n 40; exit 1; fi
ADP_FUSE_FIRST=1 timeout 300 python tools/layer_profile.py 1024 16 bf16 > gpurun_out/layers_fc.txt 2>&1; grep -E "total|first_conv|tta_input|down1_conv2" gpurun_out/layers_fc.txt
ADP_FUSE_FIRST=0 timeout 300 python tools/layer_profile.py 1024 16 bf16 > gpurun_out/layers_nofc.txt 2>&1; grep -E "total|first_conv|down1_conv2" gpurun_out/layers_nofc.txt
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --wsi none --no-x3 --train-batch 0"
for rep in 1; do
for ff in 0 1; do
  ADP_FUSE_FIRST=$ff $B > gpurun_out/ab_first_${ff}_$rep.json 2> gpurun_out/ab.err || tail -n 3 gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open('gpurun_out/ab_first_${ff}_$rep.json'))
print('fuse_first=$ff rep $rep', round(d['value'],2), 'tiles/s', round(d['ms_per_step'],2), 'ms', d['clocks'].get('sm_mhz'), 'MHz e2e', round(d['e2e']['value'],2))
PY
done
done
