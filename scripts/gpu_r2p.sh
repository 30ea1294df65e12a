#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_conv_tc_gpu.py -q -x -k "tc3" > gpurun_out/r2p_test_tc3.log 2>&1; echo "tc3 tests rc=$?"; tail -2 gpurun_out/r2p_test_tc3.log
python scripts/bench_conv.py 2>&1 | grep -v -i warn | grep "tc3 stats" | sed 's/tc2 stats.*| tc3/tc3/'
python scripts/bench_conv.py --C 64 2>&1 | grep -v -i warn | grep "tc3 stats" | sed 's/tc2 stats.*| tc3/tc3/'
for v in "RSA_LANES=2" "RSA_LANES=3" "RSA_LANES=4"; do
env $v python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench $v rc=$?"
python -c "import json;d=json.loads(open('gpurun_out/r2p_bench.json').read().splitlines()[-1]);r=d['roofline'];print('$v',d['value'],d['ms_per_step'],r['frac'],r['conv_ms_per_step'],r['in_graph']['without_conv_launches_ms'], d['single_stream_step_ms'])"
done
