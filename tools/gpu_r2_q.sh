#!/bin/bash
# round 2, GPU call Q: unconditional single-candidate fetches, no deferral; k_plan with the parent test behind the steer arithmetic
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plan_variants.py -m gpu -q -x > gpurun_out/q_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/q_pytest.log
tail -3 gpurun_out/q_pytest.log
timeout 300 python tools/micro_run.py catalina 33554432 2>&1 | grep -v "^done" | cut -c1-110
timeout 300 python tools/micro_run.py tpt 262144 | grep -v "^done"
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'])"
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none -k regex:k_edges_arc_tpe -s 1 -c 1 -o gpurun_out/q_tpe python tools/micro_run.py catalina 8388608 > gpurun_out/q_ncu_tpe.log 2>&1
