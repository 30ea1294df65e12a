#!/bin/bash
# the driver's round-end commands on the final code: GPU tests, smoke, default bench of both arms
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
python -m pytest tests -m gpu -x -q > gpurun_out/r2f_gputest.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a gpurun_out/r2f_gputest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2f_smoke.log
python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
python scripts/bench_line.py gpurun_out/r2f_bench.json
tail -3 gpurun_out/r2f_gputest.log
