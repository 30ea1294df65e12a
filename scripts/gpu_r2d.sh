#!/bin/bash
mkdir -p gpurun_out
python scripts/iso_tc3.py 2>&1 | grep -v -i warn > gpurun_out/r2d_iso_tc3.log; cat gpurun_out/r2d_iso_tc3.log
python -m pytest tests/test_model_gpu.py tests/test_conv_tc_gpu.py tests/test_kernels_gpu.py -q -x > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"; grep -E "^\[|passed|failed|Error|assert " gpurun_out/r2d_tests.log | head -20
python scripts/hostile_diff.py serial serial hostile > gpurun_out/r2d_hostile_diff.txt 2>&1; tail -3 gpurun_out/r2d_hostile_diff.txt
for v in 1 0; do
RSA_BNR_WIDE=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_bench_wide$v.json 2> gpurun_out/r2d_bench_wide$v.err; echo "bench rc=$?"
python -c "import json;d=json.loads(open('gpurun_out/r2d_bench_wide$v.json').read().splitlines()[-1]);r=d['roofline'];print('wide$v',d['value'],d['ms_per_step'],d['launches_per_step'],r['frac'],r['conv_ms_per_step'],r['per_launch_events_ms'],r['in_graph'])"
done
