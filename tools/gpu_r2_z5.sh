#!/bin/bash
# round 2, GPU call Z5: k_plan with 224-thread CTAs (4 per SM: four staged copies of the world model instead of seven)
mkdir -p gpurun_out
b() { timeout 600 python bench.py --steps 10 --warmup 3 --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench ms/step %.3f value %.4g e2e %.4g' % (d['ms_per_step'],d['value'],d['e2e']['value']))"; }
for i in 1 2; do
echo "== 128 x 7 (default)"; b
echo "== 224 x 4"; AUVRRT_LIB=$PWD/gpurun_variants/libauvrrt_p224.so b
done
