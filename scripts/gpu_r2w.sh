#!/bin/bash
# step-level switches re-measured on top of the streaming 1x1 kernel
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
for v in "RSA_X=0" "RSA_TC3_DIRECT=1" "RSA_BNR=1" "RSA_LANES=3" "RSA_LANES=0" "RSA_WGRAD_STREAM=0"; do
env $v python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench $v rc=$?"
python -c "import json;d=json.loads(open('gpurun_out/r2w_bench.json').read().splitlines()[-1]);r=d['roofline'];print('$v',round(d['value'],1),round(d['ms_per_step'],3),round(r['frac'],4),r.get('conv_ms_per_step'))"
done
