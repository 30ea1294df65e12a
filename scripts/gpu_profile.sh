#!/bin/bash
# ncu evidence for profiles/: launch list of one eager training step + full captures of the conv kernels.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
# one eager step = ~590 launches; skip model init + 3 warm-up steps + graph-less bench prologue
RSA_CUDA_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1850 -c 620 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "launchlist rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 4 -c 2 \
  -o gpurun_out/prof_conv_tc_fwd python scripts/profile_step.py > gpurun_out/ncu_fwd.log 2>&1
echo "ncu fwd rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_wgrad -s 0 -c 2 \
  -o gpurun_out/prof_conv_tc_wgrad python scripts/profile_step.py > gpurun_out/ncu_wgrad.log 2>&1
echo "ncu wgrad rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
