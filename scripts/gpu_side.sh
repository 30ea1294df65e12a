#!/bin/bash
# A/B of the plan-level concurrency (weight-gradient side stream, branch lanes): model parity tests, then the bench.
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --timeout 300 -x > gpurun_out/test_model.log 2>&1; echo "model tests rc=$?"
tail -n 4 gpurun_out/test_model.log
for cfg in "1 2" "1 2"; do
  set -- $cfg
  RSA_WGRAD_STREAM=$1 RSA_LANES=$2 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_side.json
  python - <<PY
import json; d=json.load(open("gpurun_out/bench_side.json")); print("side=$1 lanes=$2", round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
PY
done
