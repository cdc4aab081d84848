#!/bin/bash
# round 2, call P: bias-free epilogues of the data-gradient convs: train tests + training leg of the bench (compare with r2_upsum_fusion_ab.txt)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_forward.py -q -m gpu -x > gpurun_out/tests_p.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/tests_p.log
timeout 300 python tools/train_profile.py > gpurun_out/train_profile_p.txt 2>&1; grep -E "profiled|conv_dgrad_tcgen05 " gpurun_out/train_profile_p.txt
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --wsi none --no-x3"
for rep in 1 2; do
  $B > gpurun_out/ab_p_$rep.json 2> gpurun_out/ab.err || tail -n 3 gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open('gpurun_out/ab_p_$rep.json'))
print('rep $rep train', round(d['train']['ms_per_step'],3), 'ms', round(d['train']['tiles_per_s'],1), 'tiles/s | infer', round(d['value'],2), d['clocks'].get('sm_mhz'))
PY
done
