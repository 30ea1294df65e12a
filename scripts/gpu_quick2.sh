#!/bin/bash
# build, thin-conv tests, model parity (bf16), conv microbench, step profile, short bench
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -q -m gpu --timeout 120 -k "tc3" -x > gpurun_out/test_tc3.log 2>&1; echo "tc3 rc=$?"
timeout 1500 python -m pytest tests/test_model_gpu.py -q -m gpu --timeout 900 -k bf16 -x > gpurun_out/test_model.log 2>&1; echo "model(bf16) rc=$?"
timeout 300 python scripts/bench_conv.py > gpurun_out/bench_conv.log 2>&1; tail -8 gpurun_out/bench_conv.log
timeout 600 python scripts/profile_step.py --detail > gpurun_out/profile_step.log 2>&1; echo "profile rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -n 5 gpurun_out/test_tc3.log; tail -n 15 gpurun_out/test_model.log; head -24 gpurun_out/step_breakdown.txt; tail -n 2 gpurun_out/bench.log | cut -c1-600
