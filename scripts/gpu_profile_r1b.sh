#!/bin/bash
# ncu evidence for profiles/ (round 1, second half): launch list of one eager step, DRAM traffic of every 3x3 conv launch,
# full captures of the thin-layer kernels.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
RSA_CUDA_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1760 -c 700 --csv \
  --log-file gpurun_out/launches_r1b.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "launchlist rc=$?"
RSA_CUDA_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum \
  --clock-control none -k regex:conv_tc -s 603 -c 201 --csv --log-file gpurun_out/conv_traffic_r1b.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench2.log 2>&1
echo "conv traffic rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc3_kernel -s 0 -c 12 -o gpurun_out/prof_r1b_tc3_fwd python scripts/profile_step.py > gpurun_out/ncu_a.log 2>&1; echo "ncu tc3 fwd rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc3_wgrad -s 0 -c 2 -o gpurun_out/prof_r1b_tc3_wgrad python scripts/profile_step.py > gpurun_out/ncu_b.log 2>&1; echo "ncu tc3 wgrad rc=$?"
