#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:k_plan_tpt -s 1 -c 1 -o gpurun_out/j_tpt python tools/micro_run.py tpt 262144 > gpurun_out/j_ncu_tpt.log 2>&1
timeout 300 python tools/micro_run.py tpt 262144 | grep -v "^done"
ls -la gpurun_out
