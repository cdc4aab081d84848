#!/bin/bash
# round 2, call N: A/B of the mbarrier wait with a suspend-time hint (fewer spin instructions under the power cap)
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --wsi none --no-x3"
for rep in 1 2; do
for h in 0 500 4000; do
  if [ $h -eq 0 ]; then python adipose_tissue-unet_b200/build.py --force > /dev/null; else python adipose_tissue-unet_b200/build.py --force -DADP_MBAR_HINT_NS=$h > /dev/null; fi
  $B > gpurun_out/ab_hint_${h}_$rep.json 2> gpurun_out/ab.err || tail -n 3 gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open('gpurun_out/ab_hint_${h}_$rep.json'))
print('hint=$h rep $rep', round(d['value'],2), 'tiles/s', round(d['ms_per_step'],2), 'ms', d['clocks'].get('sm_mhz'), 'MHz | train', round(d['train']['ms_per_step'],2), 'ms | fc', round(d['first_conv_fusion']['tiles_per_s'],2))
PY
done
done
python adipose_tissue-unet_b200/build.py --force > /dev/null
