#!/bin/bash
# round 2, GPU call Z4: thread-per-tree planner with the world model read through L1 instead of staged per CTA
mkdir -p gpurun_out
echo "== staged (default)"; timeout 300 python tools/micro_run.py tpt 262144 | grep -v "^done"
echo "== not staged"; AUVRRT_TPT_STAGE_KB=0 timeout 300 python tools/micro_run.py tpt 262144 | grep -v "^done"
