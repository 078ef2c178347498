#!/bin/bash
# round 2, GPU call K: classification-grid resolution sweep (thread-per-edge kernel and the planner)
mkdir -p gpurun_out
for gc in 16384 65536 262144 1048576; do
  AUVRRT_GRID_CELLS=$gc timeout 300 python tools/micro_run.py catalina 33554432 2>&1 | grep -v "^done" | sed "s/^/[cells $gc] /" >> gpurun_out/k_grid.log
  AUVRRT_GRID_CELLS=$gc timeout 300 python tools/micro_run.py catalina-nocost 33554432 2>&1 | grep -v "^done" | sed "s/^/[cells $gc] /" >> gpurun_out/k_grid.log
  AUVRRT_GRID_CELLS=$gc timeout 300 python bench.py --steps 5 --warmup 3 --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[cells $gc] k_plan ms', d['ms_per_step'], d['value'])" >> gpurun_out/k_grid.log
done
cut -c1-160 gpurun_out/k_grid.log
