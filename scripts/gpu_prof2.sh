#!/bin/bash
# per-launch step detail + full ncu captures of the persistent conv kernel (C=32 / C=64 launches) and the wgrad kernel
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python scripts/profile_step.py --detail > gpurun_out/profile_step.log 2>&1; echo "profile rc=$?"
cp gpurun_out/step_breakdown.txt gpurun_out/step_detail.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc2_kernel -s 0 -c 11 -o gpurun_out/prof_tc2_fwd python scripts/profile_step.py > gpurun_out/ncu_tc2.log 2>&1; echo "ncu tc2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_wgrad -s 0 -c 2 -o gpurun_out/prof_wgrad2 python scripts/profile_step.py > gpurun_out/ncu_wg2.log 2>&1; echo "ncu wgrad rc=$?"
