#!/bin/bash
# Multi-GPU check: DP bench at N=1 and N=2 (torchrun, NCCL), short runs.
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1; echo "n1 rc=$?"
NG=$(nvidia-smi -L | wc -l)
for n in 2 4 8; do
  if [ "$NG" -ge "$n" ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.log 2>&1; echo "n$n rc=$?"
  fi
done
for f in gpurun_out/bench_n*.log; do echo "== $f"; tail -n 3 $f | cut -c1-600; done
