#!/bin/bash
# role timers of the tcgen05 conv pipeline (ADP_TC_DEBUG & 16), optionally combined with the ablation switches
mkdir -p gpurun_out
python adipose_tissue-unet_b200/build.py --force --debug > /dev/null   # the switches exist only in a debug build
for d in ${@:-16}; do
  echo "=== ADP_TC_DEBUG=$d"
  ADP_TC_DEBUG=$d timeout 120 python tools/layer_profile.py 1024 16 bf16 2>&1 | grep -E "tc-timers" | tail -20
done 2>&1 | tee gpurun_out/tc_timers.txt

python adipose_tissue-unet_b200/build.py --force > /dev/null   # back to the production library
