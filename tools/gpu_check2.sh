#!/bin/bash
# 2-GPU box: parity tests of the current build, per-layer profile, WSI N=1 vs N=2 (NCCL boundary exchange)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_train.py tests/test_gpu_post.py tests/test_gpu_wsi.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/layer_profile.py 1024 16 bf16 2>&1 | tee gpurun_out/layers.txt
bash tools/tc_timers.sh 16 | grep -E "down1_conv2|up1_conv2|up2_conv2|up1_conv3|dilate3"
timeout 300 python tools/wsi_full.py --size 8192 --overlap 0.5 2>/dev/null | grep '^{' | tee gpurun_out/wsi8k_n1.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/wsi_full.py --size 8192 --overlap 0.5 2>gpurun_out/wsi8k_n2.err | grep '^{' | tee gpurun_out/wsi8k_n2.json
tail -n 5 gpurun_out/wsi8k_n2.err
