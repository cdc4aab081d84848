#!/bin/bash
# timing experiments on the tcgen05 conv (ADP_TC_DEBUG switches of conv_tc.cuh; results are wrong by design)
mkdir -p gpurun_out
python adipose_tissue-unet_b200/build.py --force --debug > /dev/null   # the switches exist only in a debug build
for d in 0 1 2 8 10 4 6 14 15; do
  echo "=== ADP_TC_DEBUG=$d"
  ADP_TC_DEBUG=$d timeout 120 python tools/layer_profile.py 1024 16 bf16 2>&1 | grep -E "total|down1_conv2|down2_conv1|down2_conv2|up2_conv2|up1_conv1|up1_conv2|up1_conv3|up3_conv2|dilate3"
done 2>&1 | tee gpurun_out/tc_experiments.txt

python adipose_tissue-unet_b200/build.py --force > /dev/null   # back to the production library
