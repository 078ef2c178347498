#!/bin/bash
# round 2, GPU call Y3: all-pairs arc variant with two waypoints per pass over the circle table
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "edges" > gpurun_out/y_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/y_pytest.log
tail -3 gpurun_out/y_pytest.log
for mb in 3 2 4; do
echo "== minb $mb"
AUVRRT_TPE_MINB=$mb timeout 300 python tools/micro_run.py catalina-allpairs 33554432 2>&1 | grep -v "^done" | cut -c1-110
AUVRRT_TPE_MINB=$mb timeout 300 python tools/micro_run.py config4-allpairs 4194304 2>&1 | grep -v "^done" | cut -c1-110
done
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none -k regex:k_edges_arc_tpe -s 1 -c 1 -o gpurun_out/y3_tpe_ap4 python tools/micro_run.py config4-allpairs 2097152 > gpurun_out/y3_ncu_tpe_ap4.log 2>&1
