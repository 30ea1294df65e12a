#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bn_apply_kernel -s 190 -c 2 -o gpurun_out/prof_bn_apply python scripts/profile_step.py > gpurun_out/ncu_bn1.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bn_bwd_reduce_multi -s 189 -c 3 -o gpurun_out/prof_bn_red python scripts/profile_step.py > gpurun_out/ncu_bn2.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bn_bwd_apply_multi -s 189 -c 2 -o gpurun_out/prof_bn_app python scripts/profile_step.py > gpurun_out/ncu_bn3.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/ncu_bn1.log
