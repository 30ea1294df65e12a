#!/bin/bash
# optimizer + weight refresh of the deep levels overlapped with the backward of the shallow ones
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
for v in "RSA_EARLY_OPT=0" "RSA_EARLY_OPT=1"; do
env $v python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2e2_bench.json 2> gpurun_out/r2e2_bench.err; echo "bench $v rc=$?"; tail -2 gpurun_out/r2e2_bench.err
python scripts/bench_line.py gpurun_out/r2e2_bench.json
done
timeout 1200 python -m pytest tests/test_model_gpu.py -x -q > gpurun_out/r2e2_test_model.log 2>&1; echo "model tests rc=$?"; tail -3 gpurun_out/r2e2_test_model.log
