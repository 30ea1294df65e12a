#!/bin/bash
# round 2, run A: full GPU test-suite, every bench config with its CPU arm, ncu counters of the conv / HBM-bound launches
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_gpu.txt 2>&1
python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2a_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_gputest.log
tail -5 gpurun_out/r2a_gputest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err; echo "bench c2 rc=$?"
for c in 1 3 5; do
  python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r2a_bench_c$c.json 2> gpurun_out/r2a_bench_c$c.err; echo "bench c$c rc=$?"
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_ref_c2.json 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --profile-from-start off --clock-control none --metrics $M --csv --log-file gpurun_out/r2a_ncu_conv.csv python scripts/ncu_ops.py conv > gpurun_out/r2a_ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 600 ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r2a_ncu_hbm.csv python scripts/ncu_ops.py hbm > gpurun_out/r2a_ncu_hbm.log 2>&1; echo "ncu hbm rc=$?"
python scripts/bench_conv.py > gpurun_out/r2a_bench_conv.log 2>&1; echo "bench_conv rc=$?"
head -c 1500 gpurun_out/r2a_bench_c2.json
