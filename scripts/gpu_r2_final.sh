#!/bin/bash
# round 2 rehearsal of the driver's round-end commands plus the evidence under profiles/:
# GPU tests, smoke, every bench config with its CPU arm, ncu counters of the convolution / bandwidth-bound launches,
# one full ncu capture of the thin-layer kernel, the ncu launch list and the CUPTI per-launch trace of a step.
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
T=${TAG:-r2f}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_gpu.txt 2>&1
python -m pytest tests -m gpu -q --durations=5 > gpurun_out/${T}_gputest.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a gpurun_out/${T}_gputest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; echo "bench c2 rc=$?"
for c in 1 3 5; do
  python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/${T}_bench_c$c.json 2> gpurun_out/${T}_bench_c$c.err; echo "bench c$c rc=$?"
done
for c in 2 1 3 5; do
  python bench.py --impl reference --config $c --steps 3 --warmup 1 > gpurun_out/${T}_ref_c$c.json 2>/dev/null; echo "ref c$c rc=$?"
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --profile-from-start off --clock-control none --metrics $M --csv --log-file gpurun_out/${T}_ncu_conv.csv python scripts/ncu_ops.py conv > gpurun_out/${T}_ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 600 ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/${T}_ncu_hbm.csv python scripts/ncu_ops.py hbm > gpurun_out/${T}_ncu_hbm.log 2>&1; echo "ncu hbm rc=$?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc3_kernel -c 4 -o gpurun_out/${T}_prof_tc3 python scripts/ncu_ops.py conv > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"
RSA_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 700 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
RSA_WGRAD_STREAM=0 RSA_LANES=0 timeout 300 python scripts/trace_launches.py > gpurun_out/${T}_trace.log 2>&1; cp gpurun_out/trace_launches.txt gpurun_out/${T}_step_launch_trace.txt; echo "trace rc=$?"
python scripts/bench_pw.py 2>&1 | grep -v -i warn > gpurun_out/${T}_bench_pw.log
python scripts/bench_wide.py 2>&1 | grep -v -i warn > gpurun_out/${T}_bench_wide.log
python scripts/bench_conv.py 2>&1 | grep -v -i warn > gpurun_out/${T}_bench_conv.log
python scripts/bench_conv.py --C 64 2>&1 | grep -v -i warn >> gpurun_out/${T}_bench_conv.log
tail -4 gpurun_out/${T}_gputest.log; head -c 600 gpurun_out/${T}_bench_c2.json
