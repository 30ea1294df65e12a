#!/bin/bash
# streaming 1x1 kernel (pw_stream.cu): parity tests, then the step with and without it
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -x -q -k "pw_stream or pointwise or k_base" > gpurun_out/r2r_test_pw.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2r_test_pw.log
RSA_PW_STREAM=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2r_bench_off.json 2> gpurun_out/r2r_bench_off.err; echo "off rc=$?"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2r_bench_on.json 2> gpurun_out/r2r_bench_on.err; echo "on rc=$?"
python - <<'PY'
import json
for f in ("off", "on"):
    try:
        d = json.loads(open(f"gpurun_out/r2r_bench_{f}.json").read().strip().splitlines()[-1]); print(f, round(d["value"], 1), round(d["ms_per_step"], 3), round(d["e2e"]["value"], 1))
    except Exception as e: print(f, "ERR", e)
PY
RSA_WGRAD_STREAM=0 RSA_LANES=0 timeout 600 python scripts/trace_launches.py > gpurun_out/r2r_trace.log 2>&1; echo "trace rc=$?"; cp gpurun_out/trace_launches.txt gpurun_out/r2r_step_launch_trace.txt 2>/dev/null
