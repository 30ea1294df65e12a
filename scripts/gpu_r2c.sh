#!/bin/bash
# round 2, run C: which pipeline bounds conv_tc3?  RSA_TC3_DEBUG isolates TMA loads / MMAs / epilogue / TMA stores
mkdir -p gpurun_out
for dbg in 0 1 2 3 4 8 10 6 5 7; do
  echo "== RSA_TC3_DEBUG=$dbg" >> gpurun_out/r2c_trace_tc3.log
  RSA_TC3_DEBUG=$dbg python scripts/trace_tc3.py 2>&1 | grep -v -i warn >> gpurun_out/r2c_trace_tc3.log
done
cat gpurun_out/r2c_trace_tc3.log | cut -c1-400
python -m pytest tests/test_model_gpu.py -q -k "benchmarked or converges" > gpurun_out/r2c_test_model.log 2>&1; echo "model tests rc=$?"; grep -E "^\[|passed|failed|Error|assert" gpurun_out/r2c_test_model.log | head -20
python scripts/hostile_diff.py serial serial hostile > gpurun_out/r2c_hostile_diff.txt 2>&1; tail -3 gpurun_out/r2c_hostile_diff.txt
