#!/bin/bash
# round 2, call K: whole-bench A/B under the power cap: first-conv fusion on/off (inference leg), dropout fusion on/off (training leg)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --wsi none --no-x3"
for rep in 1 2; do
for ff in 0 1; do
  ADP_FUSE_FIRST=$ff $B --train-batch 0 > gpurun_out/ab_first_${ff}_$rep.json 2> gpurun_out/ab.err || tail -n 3 gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open('gpurun_out/ab_first_${ff}_$rep.json'))
print('fuse_first=$ff rep $rep', round(d['value'],2), 'tiles/s', round(d['ms_per_step'],2), 'ms', d['clocks'].get('sm_mhz'), 'MHz e2e', round(d['e2e']['value'],2))
PY
done
done
for fd in 0 1; do
  ADP_FUSE_DROPOUT=$fd $B > gpurun_out/ab_drop_$fd.json 2> gpurun_out/ab.err || tail -n 3 gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open('gpurun_out/ab_drop_$fd.json'))
print('fuse_dropout=$fd', {k:d['train'][k] for k in ('tiles_per_s','ms_per_step')})
PY
done
