#!/bin/bash
# round 2, GPU calls W2..W4: end-of-round captures, at most 64 MiB of reports per call (usage: gpu_r2_w2.sh a|b|c)
mkdir -p gpurun_out
case "$1" in
a)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/w_launches.csv python bench.py --steps 2 --warmup 1 --micro-edges 16777216 --config5-queries 65536 > gpurun_out/w_launches_bench.log 2>&1
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none -k regex:k_edges_arc_tpe -s 1 -c 1 -o gpurun_out/w_tpe_ap python tools/micro_run.py catalina-allpairs 8388608 > gpurun_out/w_ncu_tpe_ap.log 2>&1
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none -k regex:k_edges_arc_tpe -s 1 -c 1 -o gpurun_out/w_tpe_ap4 python tools/micro_run.py config4-allpairs 2097152 > gpurun_out/w_ncu_tpe_ap4.log 2>&1
;;
b)
AUVRRT_EDGES_BRUTE=1 timeout 600 ncu --section SpeedOfLight --section ComputeWorkloadAnalysis --section WarpStateStats --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis --section SchedulerStats --section InstructionStats --clock-control none -k regex:k_edges_dubins -s 1 -c 1 -o gpurun_out/w_dubins python tools/micro_run.py edges 16777216 > gpurun_out/w_ncu_dubins.log 2>&1
timeout 600 ncu --section SpeedOfLight --section ComputeWorkloadAnalysis --section WarpStateStats --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis --section SchedulerStats --section InstructionStats --clock-control none -k regex:k_plan -s 1 -c 1 -o gpurun_out/w_mode1 python tools/micro_run.py mode1 4096 > gpurun_out/w_ncu_mode1.log 2>&1
;;
c)
timeout 600 ncu --section SpeedOfLight --section ComputeWorkloadAnalysis --section WarpStateStats --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis --section SchedulerStats --section InstructionStats --clock-control none -k regex:k_nn_partial -s 1 -c 1 -o gpurun_out/w_nn python tools/micro_run.py nn 134217728 > gpurun_out/w_ncu_nn.log 2>&1
;;
esac
ls -la gpurun_out/
