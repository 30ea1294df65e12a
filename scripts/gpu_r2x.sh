#!/bin/bash
# final code on all GPUs of the box: data-parallel step (two overlapped all-reduce ranges) and the sharded scene inference
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
NG=$(nvidia-smi -L | wc -l); echo "GPUs: $NG"
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2x_bench_n1.json 2> gpurun_out/r2x_bench_n1.err; echo "n1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 20 --warmup 5 > gpurun_out/r2x_bench_n${NG}.json 2> gpurun_out/r2x_bench_n${NG}.err; echo "n$NG rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --config 5 --steps 5 --warmup 3 > gpurun_out/r2x_bench_c5_n$NG.json 2> gpurun_out/r2x_bench_c5_n$NG.err; echo "c5 n$NG rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2x_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],3), round(d["e2e"]["value"],1))
    except Exception as e: print(f, "ERR", e)
PY
