#!/bin/bash
# round 2, GPU call T: thread-per-tree planner staging experiments; captures of the shipped k_plan / thread-per-edge / Dubins kernels
mkdir -p gpurun_out
echo "== tpt default"; timeout 300 python tools/micro_run.py tpt 262144 | grep -v "^done"
echo "== tpt grid 4096 cells"; AUVRRT_GRID_CELLS=4096 timeout 300 python tools/micro_run.py tpt 262144 | grep -v "^done"
echo "== tpt probs staged (64 KB budget)"; AUVRRT_TPT_STAGE_KB=64 timeout 300 python tools/micro_run.py tpt 262144 | grep -v "^done"
echo "== tpt grid 1024 cells"; AUVRRT_GRID_CELLS=1024 timeout 300 python tools/micro_run.py tpt 262144 | grep -v "^done"
echo "== k_plan grid 4096 cells"; AUVRRT_GRID_CELLS=4096 timeout 600 python bench.py --steps 10 --warmup 3 --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'])"
echo "== k_plan default"; timeout 600 python bench.py --steps 10 --warmup 3 --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'])"
timeout 600 ncu --set full --clock-control none -k regex:k_plan -s 2 -c 1 -o gpurun_out/t_plan python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/t_ncu_plan.log 2>&1
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none -k regex:k_edges_arc_tpe -s 1 -c 1 -o gpurun_out/t_tpe python tools/micro_run.py catalina 8388608 > gpurun_out/t_ncu_tpe.log 2>&1
AUVRRT_EDGES_BRUTE=1 timeout 600 ncu --set full --clock-control none -k regex:k_edges_dubins -s 1 -c 1 -o gpurun_out/t_dubins python tools/micro_run.py edges 4194304 > gpurun_out/t_ncu_dubins.log 2>&1
ls -la gpurun_out/*.ncu-rep
