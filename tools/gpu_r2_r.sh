#!/bin/bash
# round 2, GPU call R: thread-per-tree planner with two slots per thread (paired groups) against one; TPE micro-optimisations
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_gpu_plan_variants.py -m gpu -q -x > gpurun_out/r_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r_pytest.log
tail -3 gpurun_out/r_pytest.log
echo "== spt2"; timeout 300 python tools/micro_run.py tpt 262144 | grep -v "^done"
echo "== spt1"; AUVRRT_LIB=$PWD/gpurun_variants/libauvrrt_spt1.so timeout 300 python tools/micro_run.py tpt 262144 | grep -v "^done"
echo "== spt2 524288"; timeout 300 python tools/micro_run.py tpt 524288 | grep -v "^done"
timeout 300 python tools/micro_run.py catalina 33554432 2>&1 | grep -v "^done" | cut -c1-110
timeout 300 python tools/micro_run.py catalina-nocost 33554432 2>&1 | grep -v "^done" | cut -c1-110
timeout 300 python tools/micro_run.py catalina-allpairs 33554432 2>&1 | grep -v "^done" | cut -c1-110
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none -k regex:k_plan_tpt -c 1 -o gpurun_out/r_tpt python tools/micro_run.py tpt 262144 > gpurun_out/r_ncu_tpt.log 2>&1
