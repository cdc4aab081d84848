#!/bin/bash
# round 2, call J: dropout fused into the conv epilogue: train tests, train profile fused vs unfused, short bench of the training leg
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_train.py -q -m gpu -x > gpurun_out/tests_train.log 2>&1; echo "train tests rc=$?"; tail -n 5 gpurun_out/tests_train.log
timeout 300 python tools/train_profile.py > gpurun_out/train_profile_fused.txt 2>&1; head -n 12 gpurun_out/train_profile_fused.txt
ADP_FUSE_DROPOUT=0 timeout 300 python tools/train_profile.py > gpurun_out/train_profile_unfused.txt 2>&1; head -n 3 gpurun_out/train_profile_unfused.txt
