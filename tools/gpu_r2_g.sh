#!/bin/bash
# round 2, GPU call G: mode 3 (Dubins-RRT best parent) tests, FFMA2 Dubins kernel with shared-memory waypoints
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_plan_variants.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/g_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/g_pytest.log
tail -30 gpurun_out/g_pytest.log
for mb in 2 3; do
  AUVRRT_EDGES_BRUTE=1 AUVRRT_ED_MINB=$mb timeout 300 python tools/micro_run.py edges 50000000 2>&1 | grep -v "^done" >> gpurun_out/g_dubins.log
done
cat gpurun_out/g_dubins.log
