#!/bin/bash
# round 2, GPU call L: packed all-pairs habitats/polygon, tpt non-empty-bin bits: tests + numbers
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plan_variants.py tests/test_gpu_multi.py -m gpu -q > gpurun_out/l_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/l_pytest.log
tail -6 gpurun_out/l_pytest.log
timeout 300 python tools/micro_run.py tpt 262144 | grep -v "^done"
for v in catalina catalina-allpairs; do timeout 300 python tools/micro_run.py $v 33554432 2>&1 | grep -v "^done" | cut -c1-150; done
