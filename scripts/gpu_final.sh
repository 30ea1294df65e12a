#!/bin/bash
# round-end rehearsal: what the driver runs (GPU tests, smoke, both bench arms) + refreshed ncu launch list
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/test_all_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "bench reference rc=$?"
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"
RSA_CUDA_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1760 -c 700 --csv \
  --log-file gpurun_out/launches_r1d.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "launchlist rc=$?"
tail -n 4 gpurun_out/test_all_gpu.log; tail -n 2 gpurun_out/smoke.log; tail -n 1 gpurun_out/bench_reference.log | cut -c1-700; tail -n 1 gpurun_out/bench_default.log | cut -c1-2200
