#!/bin/bash
# N-GPU verification: bench line, both full-size WSI configs, data-parallel parity
N=${1:-8}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
timeout 300 $TR --master-port 29518 tools/wsi_full.py --size 32768 --overlap 0.5 2> gpurun_out/wsi32k_n$N.err | grep '^{' > gpurun_out/wsi32k_n$N.json; echo "wsi32k rc=$?"
timeout 300 $TR --master-port 29519 tools/wsi_full.py --size 16384 --overlap 0.75 2> gpurun_out/wsi16k_n$N.err | grep '^{' > gpurun_out/wsi16k_n$N.json; echo "wsi16k rc=$?"
timeout 300 $TR --master-port 29520 tools/dp_train_check.py --precision bf16 --size 256 --batch 2 2> gpurun_out/dp_n$N.err | grep '^{' > gpurun_out/dp_n$N.json; echo "dp rc=$?"
python - <<PY
import json
for f in ("bench_n$N","wsi32k_n$N","wsi16k_n$N","dp_n$N"):
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
tail -n 3 gpurun_out/*_n$N.err | tail -n 30
