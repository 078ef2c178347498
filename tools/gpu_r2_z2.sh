#!/bin/bash
# round 2, GPU call Z2: full GPU suite on the slim-state thread-per-tree planner; its capture; config 5 at 1 GPU
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/z2_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/z2_pytest.log
tail -3 gpurun_out/z2_pytest.log
timeout 300 python tools/micro_run.py tpt 262144 | grep -v "^done"
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none -k regex:k_plan_tpt -c 1 -o gpurun_out/z2_tpt python tools/micro_run.py tpt 262144 > gpurun_out/z2_ncu_tpt.log 2>&1
