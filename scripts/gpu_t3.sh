#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -q -m gpu --timeout 120 -k "tc3" -x > gpurun_out/test_tc3.log 2>&1; echo "tc3 rc=$?"
tail -n 25 gpurun_out/test_tc3.log
timeout 300 python scripts/bench_conv.py --C 64 > gpurun_out/bench_conv64.log 2>&1; tail -9 gpurun_out/bench_conv64.log; timeout 300 python scripts/bench_conv.py > gpurun_out/bench_conv.log 2>&1; tail -10 gpurun_out/bench_conv.log
