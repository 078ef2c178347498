#!/bin/bash
# round 2, GPU call Z6: k_plan CTA shapes 224 x 4 (default) and 448 x 2; planner tests on the default
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plan_variants.py tests/test_gpu_dropin.py tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/z6_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/z6_pytest.log
tail -3 gpurun_out/z6_pytest.log
b() { timeout 600 python bench.py --steps 10 --warmup 3 --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench ms/step %.3f value %.4g e2e %.4g' % (d['ms_per_step'],d['value'],d['e2e']['value']))"; }
for i in 1 2; do
echo "== 224 x 4 (default)"; b
echo "== 448 x 2"; AUVRRT_LIB=$PWD/gpurun_variants/libauvrrt_p448.so b
done
AUVRRT_LIB=$PWD/gpurun_variants/libauvrrt_p448.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "exploring or groups" 2>&1 | tail -2
