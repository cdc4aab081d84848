#!/bin/bash
# round 2, call E: per-layer forward profile + training-step profile of the no-debug build; CLI test re-run
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 300 python tools/layer_profile.py 1024 16 bf16 > gpurun_out/layers.txt 2>&1; cat gpurun_out/layers.txt
timeout 300 python tools/train_profile.py > gpurun_out/train_profile.txt 2>&1; cat gpurun_out/train_profile.txt
timeout 900 python -m pytest tests/test_gpu_cli.py -q -m gpu -s 2>&1 | grep -E "recon nvJPEG|passed|failed|Error" | tail -5
