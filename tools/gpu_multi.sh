#!/bin/bash
# multi-GPU checks: NCCL data-parallel training parity, DP test file, bench at N GPUs
N=${1:-2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_train_dp.py tests/test_gpu_wsi.py -x -q -m gpu -s > gpurun_out/tests_multi.log 2>&1; echo "tests_multi rc=$?"
tail -n 12 gpurun_out/tests_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 --wsi-reps 1 --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -n 5 gpurun_out/bench_n$N.err
python - <<PY
import json
for l in open('gpurun_out/bench_n$N.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches','clocks')}); print(d['e2e']); print(d.get('train')); print(json.dumps(d.get('wsi'), indent=1)[:3000])
PY
