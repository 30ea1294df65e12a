#!/bin/bash
# Quick GPU iteration: build, kernel tests, TC tests, step profile.
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 300 -k "igemm or bn or thin or stem or head" > gpurun_out/test_kernels.log 2>&1; echo "kernels rc=$?"
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -q -m gpu --timeout 120 > gpurun_out/test_conv_tc.log 2>&1; echo "conv_tc rc=$?"
timeout 600 python scripts/profile_step.py --detail > gpurun_out/profile_step.log 2>&1; echo "profile rc=$?"
for f in test_kernels test_conv_tc; do echo "== $f"; tail -n 15 gpurun_out/$f.log; done
head -30 gpurun_out/step_breakdown.txt
