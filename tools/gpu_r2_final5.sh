#!/bin/bash
# round 2, end of round (fifth pass): thread-per-tree capture and the launch list of the bench on the last build
mkdir -p gpurun_out
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none -k regex:k_plan_tpt -c 1 -o gpurun_out/g_tpt python tools/micro_run.py tpt 262144 > gpurun_out/g_ncu_tpt.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/g_launches.csv python bench.py --steps 2 --warmup 1 --micro-edges 16777216 --config5-queries 65536 > gpurun_out/g_launches_bench.log 2>&1
ls -la gpurun_out | tail -4
