#!/bin/bash
# round 2, call R: resident weights with a 3-stage floor (adds up1_conv1: 147 KB of weights, 3 activation stages)
mkdir -p gpurun_out
for m in 4 3 4 3; do
  ADP_BRES_MIN_STAGES=$m timeout 300 python tools/layer_profile.py > gpurun_out/layers_bres_min$m.txt 2>&1; head -n 1 gpurun_out/layers_bres_min$m.txt; grep -E "up1_conv1|up1_conv2 " gpurun_out/layers_bres_min$m.txt
done
