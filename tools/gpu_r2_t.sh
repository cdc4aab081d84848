#!/bin/bash
# round 2, call T (re-entry session): full GPU suite of HEAD (incl. the hand-derived known-answer tests), smoke(), and an A/B of the
# number of forwards per launch (bench.py --max-forwards 16 / 32 / 64: fewer, longer launches of the persistent conv kernel)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -n 20 gpurun_out/build.log; exit 1; }
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/tests_t.log 2>&1; echo "tests rc=$?"; tail -n 4 gpurun_out/tests_t.log
grep -E "known answer" gpurun_out/tests_t.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_t.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke_t.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --wsi none --no-x3 --train-batch 0"
for rep in 1 2; do
for mf in 16 32 64; do
  timeout 300 $B --max-forwards $mf > gpurun_out/ab_t_${mf}_$rep.json 2> gpurun_out/ab_t.err || tail -n 3 gpurun_out/ab_t.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/ab_t_${mf}_$rep.json'))
    print('max_forwards=$mf rep $rep infer', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), d['clocks'].get('sm_mhz'), d['counts_tp_fp_fn_tn'], 'launches', d['gpu_launches'])
except Exception as ex:
    print('max_forwards=$mf rep $rep failed', ex)
PY
done
done
nvidia-smi --query-gpu=memory.used,memory.total --format=csv | tail -n 1
