#!/bin/bash
# streaming weight gradient of the thin 1x1 convolutions
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -x -q -k "pw_wgrad" > gpurun_out/r2v_test.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2v_test.log
python scripts/bench_pw.py 2>&1 | grep -v -i warn | grep pw_wgrad | tee gpurun_out/r2v_bench_pw.txt
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"
python -c "import json;d=json.loads(open('gpurun_out/r2v_bench.json').read().splitlines()[-1]);r=d['roofline'];print(round(d['value'],1),round(d['ms_per_step'],3),round(r['frac'],4),r.get('conv_ms_per_step'))"
timeout 900 python -m pytest tests/test_model_gpu.py -x -q -k "bf16" > gpurun_out/r2v_test_model.log 2>&1; echo "model tests rc=$?"; tail -3 gpurun_out/r2v_test_model.log
