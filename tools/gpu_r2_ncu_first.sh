#!/bin/bash
# ncu --set full of first_conv_kernel (bf16, 16 forwards of 1024^2)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
CMD="python tools/layer_profile.py 1024 16 bf16"
$CMD > gpurun_out/plain_first.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:first_conv_kernel -s 2 -c 1 -f -o gpurun_out/prof_first $CMD > gpurun_out/ncu_first.log 2>&1
echo "full capture rc=$?"; tail -n 2 gpurun_out/ncu_first.log
