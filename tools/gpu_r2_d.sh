#!/bin/bash
# round 2, GPU call D: variants of the thread-per-edge hot loop
mkdir -p gpurun_out
V=auv-sim_b200/auvrrt/variants
for lib in "" $V/lib_nocursor.so $V/lib_outhab.so; do
  for v in catalina catalina-nocost; do
    AUVRRT_LIB=$lib timeout 300 python tools/micro_run.py $v 33554432 2>&1 | grep -v "^done" | sed "s|^|[$lib] |" >> gpurun_out/d_micro.log
    AUVRRT_LIB=$lib AUVRRT_TPE_MINB=4 timeout 300 python tools/micro_run.py $v 33554432 2>&1 | grep -v "^done" | sed "s|^|[$lib] |" >> gpurun_out/d_micro.log
  done
done
cat gpurun_out/d_micro.log | cut -c1-200
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_edges_arc_tpe -s 1 -c 1 -o gpurun_out/d_tpe python tools/micro_run.py catalina 8388608 > gpurun_out/d_ncu_tpe.log 2>&1
