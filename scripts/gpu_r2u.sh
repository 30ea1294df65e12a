#!/bin/bash
# conv_tc2 with two sub-tiles per CTA (MT = 2): parity, micro-benchmark at C = 128 / 256, step with and without
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -x -q -k "tc2_3x3 or tc2_fused_bn" > gpurun_out/r2u_test.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2u_test.log
for mt in 1 2; do
RSA_TC2_MT=$mt python scripts/bench_wide.py 2>&1 | grep -v -i warn | tee gpurun_out/r2u_bench_wide_mt$mt.txt
RSA_TC2_MT=$mt python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2u_bench_mt$mt.json 2> gpurun_out/r2u_bench_mt$mt.err; echo "mt$mt rc=$?"
done
python - <<'PY'
import json
for f in ("mt1", "mt2"):
    try:
        d = json.loads(open(f"gpurun_out/r2u_bench_{f}.json").read().strip().splitlines()[-1]); r = d["roofline"]
        print(f, round(d["value"], 1), round(d["ms_per_step"], 3), round(d["e2e"]["value"], 1), r["frac"], r.get("conv_ms_per_step"))
    except Exception as e: print(f, "ERR", e)
PY
