#!/bin/bash
# round 2, GPU call Y2: all-pairs arc variant, four quads in flight, register budgets
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "edges" > gpurun_out/y_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/y_pytest.log
tail -3 gpurun_out/y_pytest.log
for mb in 2 3 4; do
echo "== minb $mb"
AUVRRT_TPE_MINB=$mb timeout 300 python tools/micro_run.py catalina-allpairs 33554432 2>&1 | grep -v "^done" | cut -c1-110
AUVRRT_TPE_MINB=$mb timeout 300 python tools/micro_run.py config4-allpairs 4194304 2>&1 | grep -v "^done" | cut -c1-110
done
