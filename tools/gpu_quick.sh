#!/bin/bash
# quick check after a kernel change: training + forward parity tests, training-step profile, forward layer profile
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_forward.py tests/test_gpu_post.py -x -q -m gpu > gpurun_out/tests_quick.log 2>&1; echo "tests rc=$?"
tail -n 3 gpurun_out/tests_quick.log
timeout 300 python tools/train_profile.py > gpurun_out/train_profile.txt 2>&1; echo "train_profile rc=$?"
head -n 24 gpurun_out/train_profile.txt
timeout 300 python tools/layer_profile.py 1024 16 bf16 tta 2>&1 | grep -E "tta|total|maxpool|add6|first_conv"
