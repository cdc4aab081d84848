#!/bin/bash
# first GPU session: post-processing + fp32/simt parity, then the tcgen05 probe under timeouts
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_post.py -x -q -m gpu > gpurun_out/post.log 2>&1; echo "post rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_forward.py -x -q -m gpu -k "fp32 or simt or front_end" > gpurun_out/fwd_fp32.log 2>&1; echo "fwd_fp32 rc=$?" >> gpurun_out/summary.txt
ADP_TC_DEBUG=0 timeout 200 python tools/tc_probe.py 256 1 > gpurun_out/probe0.log 2>&1; echo "probe0 rc=$?" >> gpurun_out/summary.txt
ADP_TC_DEBUG=1 timeout 200 python tools/tc_probe.py 256 1 > gpurun_out/probe1.log 2>&1; echo "probe1 rc=$?" >> gpurun_out/summary.txt
nvidia-smi > gpurun_out/gpu_after.txt 2>&1
cat gpurun_out/summary.txt
tail -5 gpurun_out/post.log gpurun_out/fwd_fp32.log
tail -30 gpurun_out/probe0.log
tail -30 gpurun_out/probe1.log
