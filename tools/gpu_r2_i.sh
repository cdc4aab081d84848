#!/bin/bash
# round 2, call I: first conv fused into down1_conv2 (FC variant): bit-identity test, forward tests, layer profile fused vs unfused
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_forward.py -q -m gpu -x -k "fusion" -s > gpurun_out/tests_fc.log 2>&1; rc=$?; echo "fc test rc=$rc"; tail -n 15 gpurun_out/tests_fc.log
if [ $rc -ne 0 ]; then timeout 600 compute-sanitizer --tool memcheck python tools/fc_probe.py 128 > gpurun_out/sanitizer_fc.log 2>&1; grep -v "^=========     " gpurun_out/sanitizer_fc.log | head -n 40; exit 1; fi
timeout 300 python tools/layer_profile.py 1024 16 bf16 > gpurun_out/layers_fc.txt 2>&1; grep -E "total|first_conv|down1_conv2|down2_conv1" gpurun_out/layers_fc.txt
ADP_FUSE_FIRST=0 timeout 300 python tools/layer_profile.py 1024 16 bf16 > gpurun_out/layers_nofc.txt 2>&1; grep -E "total|first_conv|down1_conv2|down2_conv1" gpurun_out/layers_nofc.txt
timeout 900 python -m pytest tests/test_gpu_forward.py -q -m gpu > gpurun_out/tests_fwd.log 2>&1; echo "fwd tests rc=$?"; tail -n 5 gpurun_out/tests_fwd.log
