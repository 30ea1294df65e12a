#!/bin/bash
mkdir -p gpurun_out
for v in "RSA_LANES=2" "RSA_LANES=3" "RSA_LANES=4" "RSA_LANES=2 RSA_BNR=1"; do
env $v python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench $v rc=$?"
python -c "import json;d=json.loads(open('gpurun_out/r2n_bench.json').read().splitlines()[-1]);r=d['roofline'];print('$v',d['value'],d['ms_per_step'],d['launches_per_step'],r['frac'],r['conv_ms_per_step'],r['in_graph']['without_conv_launches_ms'])"
done
python -m pytest tests/test_model_gpu.py -q -x > gpurun_out/r2n_test_model.log 2>&1; echo "model tests rc=$?"; tail -2 gpurun_out/r2n_test_model.log
