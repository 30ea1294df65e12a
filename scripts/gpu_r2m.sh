#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_conv_tc_gpu.py -q -x -k "tc3" > gpurun_out/r2m_test_tc3.log 2>&1; echo "tc3 tests rc=$?"; tail -2 gpurun_out/r2m_test_tc3.log
RSA_TC3_DIRECT=1 python -m pytest tests/test_conv_tc_gpu.py -q -x -k "tc3" > gpurun_out/r2m_test_tc3_direct.log 2>&1; echo "tc3 tests (direct) rc=$?"; tail -2 gpurun_out/r2m_test_tc3_direct.log
for dm in 0 1; do
echo "== RSA_TC3_DIRECT=$dm" >> gpurun_out/r2m_bench_conv.log
RSA_TC3_DIRECT=$dm timeout 300 python scripts/bench_conv.py 2>&1 | grep -v -i warn | grep "tc3 stats\|fused" >> gpurun_out/r2m_bench_conv.log
RSA_TC3_DIRECT=$dm timeout 300 python scripts/bench_conv.py --C 64 2>&1 | grep -v -i warn | grep "tc3 stats" >> gpurun_out/r2m_bench_conv.log
done
cat gpurun_out/r2m_bench_conv.log | sed 's/tc2 stats.*| tc3/tc3/'
for v in "RSA_TC3_DIRECT=0" "RSA_TC3_DIRECT=1" "RSA_BNR=0"; do
env $v python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench $v rc=$?"
python -c "import json;d=json.loads(open('gpurun_out/r2m_bench.json').read().splitlines()[-1]);r=d['roofline'];print('$v',d['value'],d['ms_per_step'],d['launches_per_step'],r['frac'],r['conv_ms_per_step'],r['in_graph']['without_conv_launches_ms'])"
done
