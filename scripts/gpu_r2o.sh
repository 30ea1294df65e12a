#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_model_gpu.py -q -k "scene or argmax" > gpurun_out/r2o_test_scene.log 2>&1; echo "scene tests rc=$?"; tail -2 gpurun_out/r2o_test_scene.log
python bench.py --config 5 --steps 10 --warmup 3 > gpurun_out/r2o_bench_c5.json 2> gpurun_out/r2o_bench_c5.err; echo "bench c5 rc=$?"
python -c "import json;d=json.loads(open('gpurun_out/r2o_bench_c5.json').read().splitlines()[-1]);print('c5',d['value'],d['ms_per_step'],d['e2e'])"
python bench.py --config 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_bench_c1.json 2> gpurun_out/r2o_bench_c1.err; echo "bench c1 rc=$?"
python -c "import json;d=json.loads(open('gpurun_out/r2o_bench_c1.json').read().splitlines()[-1]);print('c1',d['value'],d['ms_per_step'],d['e2e'])"
