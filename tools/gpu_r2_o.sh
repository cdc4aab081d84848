#!/bin/bash
# round 2, call O: fused upsample backward (EPI_UPSUM): train tests, profile on/off, bench training leg A/B
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_train.py -q -m gpu -x > gpurun_out/tests_train.log 2>&1; echo "train tests rc=$?"; tail -n 4 gpurun_out/tests_train.log
ADP_FUSE_UPSUM=1 timeout 300 python tools/train_profile.py > gpurun_out/train_profile_upsum1.txt 2>&1; grep -E "profiled|upsample|conv_dgrad_tcgen05 |dgrad_tcgen05/up1_conv1" gpurun_out/train_profile_upsum1.txt
ADP_FUSE_UPSUM=0 timeout 300 python tools/train_profile.py > gpurun_out/train_profile_upsum0.txt 2>&1; grep -E "profiled|upsample|conv_dgrad_tcgen05 |dgrad_tcgen05/up1_conv1" gpurun_out/train_profile_upsum0.txt
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --wsi none --no-x3"
for rep in 1 2; do
for f in 0 1; do
  ADP_FUSE_UPSUM=$f $B > gpurun_out/ab_upsum_${f}_$rep.json 2> gpurun_out/ab.err || tail -n 3 gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open('gpurun_out/ab_upsum_${f}_$rep.json'))
print('fuse_upsum=$f rep $rep train', round(d['train']['ms_per_step'],3), 'ms', round(d['train']['tiles_per_s'],1), 'tiles/s | infer', round(d['value'],2))
PY
done
done
