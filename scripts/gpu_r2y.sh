#!/bin/bash
# thin-layer weight gradient at large dilations: band boxes instead of nine boxes per item
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -x -q -k "tc3_wgrad or benchmarked" > gpurun_out/r2y_test.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2y_test.log
python scripts/bench_conv.py 2>&1 | grep -v -i warn | grep wgrad | tee gpurun_out/r2y_bench_conv.txt
RSA_TC3_WG_BAND=0 python scripts/bench_conv.py 2>&1 | grep -v -i warn | grep wgrad | tee gpurun_out/r2y_bench_conv_off.txt
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench rc=$?"
python -c "import json;d=json.loads(open('gpurun_out/r2y_bench.json').read().splitlines()[-1]);r=d['roofline'];print(round(d['value'],1),round(d['ms_per_step'],3),round(r['frac'],4),r.get('conv_ms_per_step'))"
