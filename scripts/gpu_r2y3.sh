#!/bin/bash
# experiment: band items for the small dilations too
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
for v in "RSA_TC3_BAND=2" "RSA_TC3_BAND=2 RSA_TC3_BAND_MIN=3"; do
echo "== $v"
env $v timeout 600 python -m pytest tests/test_conv_tc_gpu.py -x -q -k "tc3_single or tc3_fused_br or benchmarked" 2>&1 | tail -2
env $v python scripts/bench_conv.py 2>&1 | grep -v -i warn | grep "tc3 stats" | sed 's/tc2 stats.*| tc3/tc3/'
env $v python scripts/bench_conv.py --C 64 2>&1 | grep -v -i warn | grep "tc3 stats" | sed 's/tc2 stats.*| tc3/tc3/'
env $v python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2y3_bench.json 2> gpurun_out/r2y3_bench.err; echo "bench $v rc=$?"
python -c "import json;d=json.loads(open('gpurun_out/r2y3_bench.json').read().splitlines()[-1]);r=d['roofline'];print('$v',round(d['value'],1),round(d['ms_per_step'],3),round(r['frac'],4),r.get('conv_ms_per_step'))"
done
