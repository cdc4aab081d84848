#!/bin/bash
# ncu --set full of the fused first-conv + down1_conv2 kernel (first conv_tc launch of the third forward chunk)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
CMD="python tools/layer_profile.py 1024 16 bf16"
$CMD > gpurun_out/plain_fc.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 40 -c 1 -f -o gpurun_out/prof_fc $CMD > gpurun_out/ncu_fc.log 2>&1
echo "full capture rc=$?"; tail -n 3 gpurun_out/ncu_fc.log; ls -la gpurun_out/prof_fc*
