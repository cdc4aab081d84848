#!/bin/bash
# 8-GPU box: bench line at N=8, both full-size WSI configs at N=2,4,8 (NCCL boundary exchange), data-parallel parity at N=8
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
run() { timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
run 8 29517 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench rc=$?"
port=29530
for N in 8 4 2; do
  port=$((port+1)); run $N $port tools/wsi_full.py --size 32768 --overlap 0.5 2> gpurun_out/wsi32k_n$N.err | grep '^{' > gpurun_out/wsi32k_n$N.json; echo "wsi32k N=$N rc=$?"
  port=$((port+1)); run $N $port tools/wsi_full.py --size 16384 --overlap 0.75 2> gpurun_out/wsi16k_n$N.err | grep '^{' > gpurun_out/wsi16k_n$N.json; echo "wsi16k N=$N rc=$?"
done
run 8 29520 tools/dp_train_check.py --precision bf16 --size 256 --batch 2 2> gpurun_out/dp_n8.err | grep '^{' > gpurun_out/dp_n8.json; echo "dp rc=$?"
python - <<PY
import json, glob
for f in ["bench_n8"] + [f"wsi32k_n{n}" for n in (2,4,8)] + [f"wsi16k_n{n}" for n in (2,4,8)] + ["dp_n8"]:
    try:
        for l in open(f"gpurun_out/{f}.json"):
            if l.startswith("{"):
                d=json.loads(l)
                if f.startswith("bench"):
                    print(f, {k:d[k] for k in ("value","n_gpus","ms_per_step")}, d["e2e"]["value"], d.get("train",{}).get("tiles_per_s"), d.get("train",{}).get("ms_per_step"), d.get("train",{}).get("collective_path"), d.get("wsi",{}).get("mpx_per_s"))
                else:
                    print(f, {k:d[k] for k in d if k in ("seconds","mpx_per_s","tiles_per_s","counts_sum_equals_pixels","rank0_phases_s","ok","loss_rel","theta_rms_ratio","replicas_identical","world")})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -n 2 gpurun_out/*_n8.err | tail -n 20
