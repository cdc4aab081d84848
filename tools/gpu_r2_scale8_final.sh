#!/bin/bash
# final code at 8 GPUs of one box: bench.py (both full-size whole-slide cases, data-parallel training step)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --wsi-reps 1 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench N=8 rc=$?"
python - <<PY
import json
for l in open("gpurun_out/bench_n8.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print({k: d[k] for k in ("value", "n_gpus", "ms_per_step")}, "e2e", d["e2e"]["value"])
        print("  train", {k: d["train"][k] for k in ("tiles_per_s", "ms_per_step", "wall_ms_per_step")})
        for k, w in d["wsi"].items():
            print("  wsi", k, {q: w[q] for q in ("seconds_reps", "mpx_per_s", "frac_of_ceiling", "counts_tp_fp_fn_tn")})
PY
tail -n 3 gpurun_out/bench_n8.err
