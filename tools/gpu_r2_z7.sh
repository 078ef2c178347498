#!/bin/bash
# round 2, GPU call Z7: k_plan 448 x 2 with the probability table staged as well
b() { timeout 600 python bench.py --steps 10 --warmup 3 --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench ms/step %.3f value %.4g e2e %.4g' % (d['ms_per_step'],d['value'],d['e2e']['value']))"; }
echo "== 448 x 2 hot part staged (default)"; b
echo "== 448 x 2 hot + probs staged"; AUVRRT_PLAN_STAGE_KB=64 b
echo "== 448 x 2 grid 4096 cells"; AUVRRT_GRID_CELLS=4096 b
echo "== 448 x 2 grid 4096 cells, probs staged"; AUVRRT_GRID_CELLS=4096 AUVRRT_PLAN_STAGE_KB=64 b
