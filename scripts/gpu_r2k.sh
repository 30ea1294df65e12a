#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_conv_tc_gpu.py -q -x -k "tc3" > gpurun_out/r2k_test_tc3.log 2>&1; echo "tc3 tests rc=$?"; tail -3 gpurun_out/r2k_test_tc3.log
timeout 300 python scripts/bench_conv.py 2>&1 | grep -v -i warn > gpurun_out/r2k_bench_conv.log; cat gpurun_out/r2k_bench_conv.log
timeout 300 python scripts/bench_conv.py --C 64 2>&1 | grep -v -i warn > gpurun_out/r2k_bench_conv64.log; cat gpurun_out/r2k_bench_conv64.log
python -m pytest tests/test_model_gpu.py -q > gpurun_out/r2k_test_model.log 2>&1; echo "model tests rc=$?"; grep -E "^\[|passed|failed|Error|assert " gpurun_out/r2k_test_model.log | head
for fb in 1 0; do
RSA_FUSE_BRANCHES=$fb python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2k_bench_fb$fb.json 2> gpurun_out/r2k_bench_fb$fb.err; echo "bench rc=$?"
python -c "import json;d=json.loads(open('gpurun_out/r2k_bench_fb$fb.json').read().splitlines()[-1]);r=d['roofline'];print('fuse_branches=$fb',d['value'],d['ms_per_step'],d['launches_per_step'],r['frac'],r['conv_ms_per_step'],r['launches'],r['in_graph']);print(r['by_kernel'])"
done
