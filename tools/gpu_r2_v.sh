#!/bin/bash
# round 2, GPU call V: full GPU test suite on the end-of-round build; k_plan, config-4 and Catalina numbers; first batch of captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/v_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/v_pytest.log
tail -4 gpurun_out/v_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench ms/step %.3f value %.4g e2e %.4g' % (d['ms_per_step'],d['value'],d['e2e']['value']))"
for sh in 256 1024; do echo "== config4 nocost threads $sh"; AUVRRT_TPE_THREADS=$sh timeout 300 python tools/micro_run.py config4-nocost 33554432 2>&1 | grep -v "^done" | cut -c1-100; done
for sh in 256 1024; do echo "== config4 cost threads $sh"; AUVRRT_TPE_THREADS=$sh timeout 300 python tools/micro_run.py config4 33554432 2>&1 | grep -v "^done" | cut -c1-100; done
timeout 300 python tools/micro_run.py catalina 33554432 2>&1 | grep -v "^done" | cut -c1-110
timeout 300 python tools/micro_run.py catalina-nocost 33554432 2>&1 | grep -v "^done" | cut -c1-110
timeout 300 python tools/micro_run.py tpt 262144 | grep -v "^done"
timeout 600 ncu --set full --clock-control none -k regex:k_plan -s 2 -c 1 -o gpurun_out/v_plan python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/v_ncu_plan.log 2>&1
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none -k regex:k_edges_arc_tpe -s 1 -c 1 -o gpurun_out/v_tpe python tools/micro_run.py catalina 8388608 > gpurun_out/v_ncu_tpe.log 2>&1
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none -k regex:k_plan_tpt -c 1 -o gpurun_out/v_tpt python tools/micro_run.py tpt 262144 > gpurun_out/v_ncu_tpt.log 2>&1
ls -la gpurun_out/*.ncu-rep
