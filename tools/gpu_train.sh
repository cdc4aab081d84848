#!/bin/bash
# training-step parity tests (+ the forward suite when FULL=1)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 1200 python -m pytest tests/test_gpu_train.py -q -m gpu -s ${KSEL:+-k "$KSEL"} > gpurun_out/tests_train.log 2>&1; echo "tests_train rc=$?"
tail -n 30 gpurun_out/tests_train.log
if [ "$FULL" == "1" ]; then
  timeout 1200 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_train.py > gpurun_out/tests.log 2>&1; echo "tests rc=$?"
  tail -n 5 gpurun_out/tests.log
fi
