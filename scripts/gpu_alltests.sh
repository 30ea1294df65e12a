#!/bin/bash
# exactly what the driver runs at round end: all GPU tests, smoke, default bench
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/test_all_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"
tail -n 8 gpurun_out/test_all_gpu.log; tail -n 3 gpurun_out/smoke.log; tail -n 1 gpurun_out/bench_default.log | cut -c1-1500
