#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_conv_tc_gpu.py -q -x -k "tc3" > gpurun_out/r2f_test_tc3.log 2>&1; echo "tc3 tests rc=$?"; tail -3 gpurun_out/r2f_test_tc3.log
for n in 16 8 32; do echo "== N=$n" >> gpurun_out/r2f_iso_tc3.log; N=$n timeout 600 python scripts/iso_tc3.py 2>&1 | grep -v -i warn | grep "d= 1" >> gpurun_out/r2f_iso_tc3.log; done; cat gpurun_out/r2f_iso_tc3.log
timeout 300 python scripts/bench_conv.py 2>&1 | grep -v -i warn > gpurun_out/r2f_bench_conv.log; cat gpurun_out/r2f_bench_conv.log
timeout 300 python scripts/bench_conv.py --C 64 2>&1 | grep -v -i warn > gpurun_out/r2f_bench_conv64.log; cat gpurun_out/r2f_bench_conv64.log
python -m pytest tests/test_model_gpu.py -q > gpurun_out/r2f_test_model.log 2>&1; echo "model tests rc=$?"; grep -E "^\[|passed|failed|Error|assert " gpurun_out/r2f_test_model.log | head
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
python -c "import json;d=json.loads(open('gpurun_out/r2f_bench.json').read().splitlines()[-1]);r=d['roofline'];print(d['value'],d['ms_per_step'],d['launches_per_step'],r['frac'],r['conv_ms_per_step'],r['per_launch_events_ms'],r['in_graph']);print(r['by_kernel'])"
