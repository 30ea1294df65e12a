#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -q -m gpu --timeout 120 -x > gpurun_out/test_conv_tc.log 2>&1; echo "conv_tc rc=$?"
tail -n 5 gpurun_out/test_conv_tc.log
timeout 600 python scripts/profile_step.py --detail > gpurun_out/profile_step.log 2>&1; echo "profile rc=$?"
head -14 gpurun_out/step_breakdown.txt
