#!/bin/bash
# thin-layer forward / data gradient at large dilations: band items
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -x -q -k "tc3" > gpurun_out/r2y2_test.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2y2_test.log
python scripts/bench_conv.py 2>&1 | grep -v -i warn | grep "tc3 stats" | sed 's/tc2 stats.*| tc3/tc3/' | tee gpurun_out/r2y2_bench_conv.txt
python scripts/bench_conv.py --C 64 2>&1 | grep -v -i warn | grep "tc3 stats" | sed 's/tc2 stats.*| tc3/tc3/' | tee -a gpurun_out/r2y2_bench_conv.txt
for v in "RSA_TC3_BAND=0" "RSA_TC3_BAND=1"; do
env $v python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2y2_bench.json 2> gpurun_out/r2y2_bench.err; echo "bench $v rc=$?"
python -c "import json;d=json.loads(open('gpurun_out/r2y2_bench.json').read().splitlines()[-1]);r=d['roofline'];print('$v',round(d['value'],1),round(d['ms_per_step'],3),round(r['frac'],4),r.get('conv_ms_per_step'))"
done
