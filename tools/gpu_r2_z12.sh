#!/bin/bash
# round 2, GPU call Z12: k_plan 896 x 1 with the world model, the probability table and the grid plane in shared memory
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/z12_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/z12_pytest.log
tail -3 gpurun_out/z12_pytest.log
b() { timeout 600 python bench.py --steps 10 --warmup 3 --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench ms/step %.3f value %.4g e2e %.4g' % (d['ms_per_step'],d['value'],d['e2e']['value']))"; }
b; b
