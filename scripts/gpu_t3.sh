#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -q -m gpu --timeout 120 -k "tc3" -x > gpurun_out/test_tc3.log 2>&1; echo "tc3 rc=$?"
tail -n 5 gpurun_out/test_tc3.log
for ew in 16 8; do
  echo "== RSA_TC3_EW=$ew"
  RSA_TC3_EW=$ew timeout 300 python scripts/bench_conv.py --C 64 2>&1 | grep "tc3 stats" | sed 's/.*| tc3/tc3/' | paste -d' ' <(printf "C=64 d=1 \nC=64 d=3 \nC=64 d=15\nC=64 d=31\n") -
  RSA_TC3_EW=$ew timeout 300 python scripts/bench_conv.py 2>&1 | grep "tc3 stats\|fused" | sed 's/.*| tc3/tc3/'
done
