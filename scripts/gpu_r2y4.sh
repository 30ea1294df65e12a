#!/bin/bash
# 64-channel band launches: weight gradient (single-row items), forward / data gradient with side inputs (no staging buffers)
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -x -q -k "tc3" > gpurun_out/r2y4_test.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2y4_test.log
python scripts/bench_conv.py 2>&1 | grep -v -i warn | sed 's/tc2 stats.*| tc3/tc3/' | tee gpurun_out/r2y4_bench_conv.txt
python scripts/bench_conv.py --C 64 2>&1 | grep -v -i warn | sed 's/tc2 stats.*| tc3/tc3/' | tee -a gpurun_out/r2y4_bench_conv.txt
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2y4_bench.json 2> gpurun_out/r2y4_bench.err; echo "bench rc=$?"
python -c "import json;d=json.loads(open('gpurun_out/r2y4_bench.json').read().splitlines()[-1]);r=d['roofline'];print(round(d['value'],1),round(d['ms_per_step'],3),round(r['frac'],4),r.get('conv_ms_per_step'))"
timeout 900 python -m pytest tests/test_model_gpu.py -x -q > gpurun_out/r2y4_test_model.log 2>&1; echo "model tests rc=$?"; tail -4 gpurun_out/r2y4_test_model.log
