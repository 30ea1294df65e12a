#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bn_bwd_reduce_multi -s 191 -c 2 -o gpurun_out/prof_bn_red python scripts/profile_step.py > gpurun_out/ncu_bn2.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bn_bwd_apply_multi -s 191 -c 2 -o gpurun_out/prof_bn_app python scripts/profile_step.py > gpurun_out/ncu_bn3.log 2>&1; echo "rc=$?"
