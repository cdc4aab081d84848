#!/bin/bash
# data-gradient twin: mask staging modes (ADP_BWD_MASK, engine.cu launch_conv_tc) - parity tests once, then one profiled step per mode
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_train.py -q -m gpu > gpurun_out/tests_train.log 2>&1; echo "tests_train rc=$?"
tail -n 3 gpurun_out/tests_train.log
for m in auto 0 1 2 3; do
  if [ "$m" == "auto" ]; then unset ADP_BWD_MASK; else export ADP_BWD_MASK=$m; fi
  timeout 300 python tools/train_profile.py > gpurun_out/train_profile_mask_$m.txt 2>&1; echo "mode $m rc=$?"
  grep -E "train step|profiled step|conv_dgrad_tcgen05  " gpurun_out/train_profile_mask_$m.txt
done
