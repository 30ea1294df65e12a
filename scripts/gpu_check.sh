#!/bin/bash
# One gpurun call: build check, kernel + model parity tests, smoke, short bench.  Logs -> gpurun_out/
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 300 > gpurun_out/test_kernels.log 2>&1
echo "kernels rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -q -m gpu --timeout 120 > gpurun_out/test_conv_tc.log 2>&1
echo "conv_tc rc=$?" >> gpurun_out/summary.txt
RSA_CONV_ENGINE=simt timeout 1500 python -m pytest tests/test_model_gpu.py -q -m gpu --timeout 900 > gpurun_out/test_model_simt.log 2>&1
echo "model(simt) rc=$?" >> gpurun_out/summary.txt
timeout 1500 python -m pytest tests/test_model_gpu.py -q -m gpu --timeout 900 -k bf16 > gpurun_out/test_model.log 2>&1
echo "model(tc,bf16) rc=$?" >> gpurun_out/summary.txt
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 600 python scripts/profile_step.py --detail > gpurun_out/profile_step.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/summary.txt
for f in test_kernels test_conv_tc test_model_simt test_model smoke profile_step bench; do echo "== $f"; tail -n 12 gpurun_out/$f.log; done
cat gpurun_out/summary.txt
