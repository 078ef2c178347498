#!/bin/bash
# round 2, GPU call P: outlined queue push
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "edges or pins or fp32_cost or shapes" > gpurun_out/p_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/p_pytest.log
tail -3 gpurun_out/p_pytest.log
timeout 300 python tools/micro_run.py catalina 33554432 2>&1 | grep -v "^done" | cut -c1-110
AUVRRT_TPE_THREADS=256 timeout 300 python tools/micro_run.py catalina 33554432 2>&1 | grep -v "^done" | cut -c1-110
timeout 300 python tools/micro_run.py catalina-nocost 33554432 2>&1 | grep -v "^done" | cut -c1-110
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none -k regex:k_edges_arc_tpe -s 1 -c 1 -o gpurun_out/p_tpe python tools/micro_run.py catalina 8388608 > gpurun_out/p_ncu_tpe.log 2>&1
