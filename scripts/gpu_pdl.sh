#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_kernels_gpu.py -q -m gpu --timeout 300 -x > gpurun_out/test_k.log 2>&1; echo "kernel tests rc=$?"
timeout 1500 python -m pytest tests/test_model_gpu.py -q -m gpu --timeout 900 -x > gpurun_out/test_model.log 2>&1; echo "model rc=$?"
tail -n 3 gpurun_out/test_k.log; tail -n 3 gpurun_out/test_model.log
for i in 1 2; do
for v in 1 0; do
RSA_PDL=$v python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('PDL=$v', round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1))"
done; done
