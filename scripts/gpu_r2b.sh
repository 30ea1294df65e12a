#!/bin/bash
# round 2, run B: two persistent conv_tc3 CTAs per SM (RSA_TC3_OCC=2) against one, kernel tests, run-to-run reproducibility
mkdir -p gpurun_out
python -m pytest tests/test_conv_tc_gpu.py -q -k "tc3" > gpurun_out/r2b_test_tc3.log 2>&1; echo "tc3 tests rc=$?"; tail -3 gpurun_out/r2b_test_tc3.log
python -m pytest tests/test_model_gpu.py -q -k "benchmarked or hostile or converges" > gpurun_out/r2b_test_model.log 2>&1; echo "model tests rc=$?"; tail -5 gpurun_out/r2b_test_model.log
for occ in 1 2 1 2; do
  echo "== RSA_TC3_OCC=$occ" >> gpurun_out/r2b_bench_conv.log
  RSA_TC3_OCC=$occ python scripts/bench_conv.py 2>&1 | grep -v -i warn >> gpurun_out/r2b_bench_conv.log
done
cat gpurun_out/r2b_bench_conv.log
for occ in 1 2; do
  RSA_TC3_OCC=$occ python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench_occ$occ.json 2> gpurun_out/r2b_bench_occ$occ.err; echo "bench occ$occ rc=$?"
  python -c "import json;d=json.loads(open('gpurun_out/r2b_bench_occ$occ.json').read().splitlines()[-1]);print('occ$occ',d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['conv_ms_per_step'])"
done
python scripts/hostile_diff.py serial serial hostile > gpurun_out/r2b_hostile_diff.txt 2>&1; tail -3 gpurun_out/r2b_hostile_diff.txt
