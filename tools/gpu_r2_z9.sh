#!/bin/bash
# round 2, GPU call Z9: thread-per-tree planner with two barriers per trip instead of four
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plan_variants.py tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/z9_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/z9_pytest.log
tail -3 gpurun_out/z9_pytest.log
timeout 300 python tools/micro_run.py tpt 262144 | grep -v "^done"
timeout 300 python tools/micro_run.py tpt 524288 | grep -v "^done"
