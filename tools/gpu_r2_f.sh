#!/bin/bash
# round 2, GPU call F: ncu captures of the current kernels
mkdir -p gpurun_out
AUVRRT_EDGES_BRUTE=1 timeout 600 ncu --set full --clock-control none -k regex:k_edges_dubins -s 1 -c 1 -o gpurun_out/f_dubins python tools/micro_run.py edges 10000000 > gpurun_out/f_ncu_dubins.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_plan -s 2 -c 1 -o gpurun_out/f_plan python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/f_ncu_plan.log 2>&1
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none -k regex:k_edges_arc_tpe -s 1 -c 1 -o gpurun_out/f_tpe python tools/micro_run.py catalina 8388608 > gpurun_out/f_ncu_tpe.log 2>&1
ls -la gpurun_out | tail -5
