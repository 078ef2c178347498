#!/bin/bash
# round 2, GPU call Z11: k_plan 896 x 1 as the default: planner tests; probability table staged as well
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plan_variants.py tests/test_gpu_dropin.py tests/test_gpu_multi.py tests/test_bench_contract.py -m gpu -q -x > gpurun_out/z11_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/z11_pytest.log
tail -3 gpurun_out/z11_pytest.log
b() { timeout 600 python bench.py --steps 10 --warmup 3 --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench ms/step %.3f value %.4g e2e %.4g' % (d['ms_per_step'],d['value'],d['e2e']['value']))"; }
echo "== 896 x 1 (default)"; b
echo "== 896 x 1, probs staged"; AUVRRT_PLAN_BUDGET_KB=64 b
echo "== 896 x 1 (default)"; b
echo "== 896 x 1, probs staged"; AUVRRT_PLAN_BUDGET_KB=64 b
