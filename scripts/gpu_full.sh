#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -q -m gpu --timeout 120 -x > gpurun_out/test_conv_tc.log 2>&1; echo "conv_tc rc=$?"
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 300 > gpurun_out/test_kernels.log 2>&1; echo "kernel tests rc=$?"
timeout 1500 python -m pytest tests/test_model_gpu.py -q -m gpu --timeout 900 -k bf16 -x > gpurun_out/test_model.log 2>&1; echo "model(bf16) rc=$?"
timeout 600 python scripts/profile_step.py --detail > gpurun_out/profile_step.log 2>&1; echo "profile rc=$?"
tail -n 3 gpurun_out/test_conv_tc.log; tail -n 3 gpurun_out/test_kernels.log; tail -n 4 gpurun_out/test_model.log; head -24 gpurun_out/step_breakdown.txt
