#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 300 -k "bn" > gpurun_out/test_kernels.log 2>&1; echo "bn tests rc=$?"
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -q -m gpu --timeout 120 -k "tc3" -x > gpurun_out/test_tc3.log 2>&1; echo "tc3 rc=$?"
timeout 600 python scripts/profile_step.py --detail > gpurun_out/profile_step.log 2>&1; echo "profile rc=$?"
tail -n 3 gpurun_out/test_kernels.log; tail -n 3 gpurun_out/test_tc3.log; head -16 gpurun_out/step_breakdown.txt
grep -A40 "non-conv launches" gpurun_out/step_breakdown.txt | grep "bn_bwd" | head -12
