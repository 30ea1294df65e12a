#!/bin/bash
# A/B of the weight-gradient side stream: model parity tests, then the bench with and without it.
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_conv_tc_gpu.py -q -m gpu --timeout 300 -x > gpurun_out/test_model.log 2>&1; echo "model+conv tests rc=$?"
tail -n 4 gpurun_out/test_model.log
for v in 1 0 1 0; do
  RSA_WGRAD_STREAM=$v python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_side$v.json
  python - <<PY
import json; d=json.load(open("gpurun_out/bench_side$v.json")); print("RSA_WGRAD_STREAM=$v", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
done
