#!/bin/bash
# round 2, call B: full GPU suite only (continue past failures), report
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 2000 python -m pytest tests -q -m gpu -s > gpurun_out/tests.log 2>&1; echo "tests rc=$?"
grep -E "configs\[1\]|1024\^2|256\^2|trained|ohem|passed|failed|FAILED|Error" gpurun_out/tests.log | tail -n 60
