#!/bin/bash
# round 2, 8-GPU box: the driver's scaling run in miniature - bench.py at N = 8 (and 4) with both full-size WSI cases and the
# data-parallel training step (bucketed NCCL all-reduce overlapped with the backward pass), DP parity check at N = 8
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
run 8 29517 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --wsi-reps 2 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench N=8 rc=$?"
run 4 29518 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline --wsi-reps 1 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "bench N=4 rc=$?"
run 8 29520 tools/dp_train_check.py --precision bf16 --size 256 --batch 2 2> gpurun_out/dp_n8.err | grep '^{' > gpurun_out/dp_n8.json; echo "dp rc=$?"
python - <<PY
import json
for f in ("bench_n8", "bench_n4"):
    for l in open(f"gpurun_out/{f}.json"):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, {k: d[k] for k in ("value", "n_gpus", "ms_per_step")}, "e2e", d["e2e"]["value"])
            print("  train", {k: d["train"][k] for k in ("tiles_per_s", "ms_per_step", "wall_ms_per_step")})
            for k, w in d["wsi"].items():
                print("  wsi", k, {q: w[q] for q in ("seconds_reps", "mpx_per_s", "frac_of_ceiling", "counts_tp_fp_fn_tn")})
print(open("gpurun_out/dp_n8.json").read())
PY
tail -n 3 gpurun_out/bench_n8.err gpurun_out/dp_n8.err | tail -n 12
