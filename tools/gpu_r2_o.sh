#!/bin/bash
# round 2, GPU call O: deferred slow cases in the thread-per-edge kernel; fresh k_plan profile
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/o_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/o_pytest.log
tail -4 gpurun_out/o_pytest.log
timeout 300 python tools/micro_run.py catalina 33554432 2>&1 | grep -v "^done" | cut -c1-110
AUVRRT_TPE_GRIDS=0 timeout 300 python tools/micro_run.py catalina 33554432 2>&1 | grep -v "^done" | cut -c1-110
AUVRRT_TPE_THREADS=512 timeout 300 python tools/micro_run.py catalina 33554432 2>&1 | grep -v "^done" | cut -c1-110
timeout 300 python tools/micro_run.py catalina-nocost 33554432 2>&1 | grep -v "^done" | cut -c1-110
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none -k regex:k_edges_arc_tpe -s 1 -c 1 -o gpurun_out/o_tpe python tools/micro_run.py catalina 8388608 > gpurun_out/o_ncu_tpe.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_plan -s 2 -c 1 -o gpurun_out/o_plan python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/o_ncu_plan.log 2>&1
ls -la gpurun_out | tail -5
