#!/bin/bash
# round 2, GPU call N: profile of the thread-per-edge kernel in its new default shape (1024 threads, grid plane in shared memory); tpt on the straight-line lookups
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tpt or thread_per_tree or edges" > gpurun_out/n_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/n_pytest.log
tail -4 gpurun_out/n_pytest.log
timeout 300 python tools/micro_run.py tpt 262144 | grep -v "^done"
timeout 300 python tools/micro_run.py catalina 33554432 2>&1 | grep -v "^done" | cut -c1-110
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none -k regex:k_edges_arc_tpe -s 1 -c 1 -o gpurun_out/n_tpe python tools/micro_run.py catalina 8388608 > gpurun_out/n_ncu_tpe.log 2>&1
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none -k regex:k_plan_tpt -c 1 -o gpurun_out/n_tpt python tools/micro_run.py tpt 262144 > gpurun_out/n_ncu_tpt.log 2>&1
ls -la gpurun_out | tail -5
