#!/bin/bash
# role timers (debug build) of down1_conv2 with the first conv fused in / not
mkdir -p gpurun_out
python adipose_tissue-unet_b200/build.py --force --debug > /dev/null
for ff in 1 0; do
  echo "=== ADP_FUSE_FIRST=$ff"
  ADP_FUSE_FIRST=$ff ADP_TC_DEBUG=16 timeout 120 python tools/layer_profile.py 1024 16 bf16 2>&1 | grep -E "tc-timers.*down1_conv2" | tail -1
done 2>&1 | tee gpurun_out/tc_timers_fc.txt
python adipose_tissue-unet_b200/build.py --force > /dev/null
