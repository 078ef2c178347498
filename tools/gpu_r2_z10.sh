#!/bin/bash
# round 2, GPU call Z10: slim per-group scratch in the planners; k_plan 448 x 2 (default) against 896 x 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/z10_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/z10_pytest.log
tail -3 gpurun_out/z10_pytest.log
b() { timeout 600 python bench.py --steps 10 --warmup 3 --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench ms/step %.3f value %.4g e2e %.4g' % (d['ms_per_step'],d['value'],d['e2e']['value']))"; }
for i in 1 2; do
echo "== 448 x 2 (default)"; b
echo "== 896 x 1"; AUVRRT_LIB=$PWD/gpurun_variants/libauvrrt_p896.so b
done
AUVRRT_LIB=$PWD/gpurun_variants/libauvrrt_p896.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "exploring or groups" 2>&1 | tail -2
