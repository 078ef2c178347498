#!/bin/bash
# round 2, GPU call E: full GPU tests; k_plan one-chunk variant and lane-group sizes; FFMA2 Dubins all-pairs kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/e_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/e_pytest.log
tail -6 gpurun_out/e_pytest.log
for g in 0 16 8; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --group $g 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('group $g', d['ms_per_step'], d['value'], d['roofline']['kernel'])" >> gpurun_out/e_plan.log
done
cat gpurun_out/e_plan.log
for mb in 1 2 3; do
  AUVRRT_EDGES_BRUTE=1 AUVRRT_ED_MINB=$mb timeout 300 python tools/micro_run.py edges 50000000 2>&1 | grep -v "^done" >> gpurun_out/e_dubins.log
done
AUVRRT_EDGES_BRUTE=0 timeout 300 python tools/micro_run.py edges 50000000 2>&1 | grep -v "^done" >> gpurun_out/e_dubins.log
cat gpurun_out/e_dubins.log
AUVRRT_EDGES_BRUTE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_edges_dubins -s 1 -c 1 -o gpurun_out/e_dubins python tools/micro_run.py edges 10000000 > gpurun_out/e_ncu_dubins.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_plan -s 2 -c 1 -o gpurun_out/e_plan python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/e_ncu_plan.log 2>&1
ls -la gpurun_out | tail -5
