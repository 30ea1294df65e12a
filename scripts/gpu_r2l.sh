#!/bin/bash
# round 2, run L (2 GPUs): equivalence of the overlapped data-parallel step, N=1 / N=2 bench with 1 and 2 overlapped ranges
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
nvidia-smi -L > gpurun_out/r2l_gpus.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/dp_check.py > gpurun_out/r2l_dp_check.log 2>&1; echo "dp_check rc=$?"; tail -6 gpurun_out/r2l_dp_check.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2l_bench_n1.json 2> gpurun_out/r2l_bench_n1.err; echo "n1 rc=$?"
for r in 2 1; do
RSA_DP_RANGES=$r timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2l_bench_n2_r$r.json 2> gpurun_out/r2l_bench_n2_r$r.err; echo "n2 ranges=$r rc=$?"
done
python - <<'PY'
import json
for f in ("r2l_bench_n1","r2l_bench_n2_r2","r2l_bench_n2_r1"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("dp_mode"))
    except Exception as e: print(f, "ERR", e)
PY
