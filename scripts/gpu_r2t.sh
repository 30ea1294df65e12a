#!/bin/bash
# streaming 1x1 kernel with the cp.async FIFO: parity, micro-benchmark, step
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -x -q -k "pw_stream or pointwise or k_base" > gpurun_out/r2t_test_pw.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2t_test_pw.log
python scripts/bench_pw.py --wgrad 0 2>&1 | grep -v -i warn | tee gpurun_out/r2t_bench_pw.txt
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2t_bench_on.json 2> gpurun_out/r2t_bench_on.err; echo "on rc=$?"
python - <<'PY'
import json
for f in ("on",):
    try:
        d = json.loads(open(f"gpurun_out/r2t_bench_{f}.json").read().strip().splitlines()[-1]); print(f, round(d["value"], 1), round(d["ms_per_step"], 3), round(d["e2e"]["value"], 1))
    except Exception as e: print(f, "ERR", e)
PY
