#!/bin/bash
# round 2, GPU call M: branch-free single-candidate collision / habitat tests; CTA shapes and the grid plane in shared memory
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plan_variants.py tests/test_gpu_dropin.py -m gpu -q -x > gpurun_out/m_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/m_pytest.log
tail -12 gpurun_out/m_pytest.log
run() { echo "== $*"; env "$@" timeout 300 python tools/micro_run.py catalina 33554432 2>&1 | grep -v "^done" | cut -c1-110; }
run A=0
run AUVRRT_TPE_MINB=3
run AUVRRT_TPE_THREADS=512
run AUVRRT_TPE_THREADS=1024
run AUVRRT_TPE_THREADS=512 AUVRRT_TPE_GRIDS=1
run AUVRRT_TPE_THREADS=1024 AUVRRT_TPE_GRIDS=1
run AUVRRT_GRID_CELLS=4096 AUVRRT_TPE_GRIDS=1
run AUVRRT_GRID_CELLS=4096 AUVRRT_TPE_GRIDS=1 AUVRRT_TPE_MINB=3
run AUVRRT_GRID_CELLS=4096
run AUVRRT_GRID_CELLS=65536 AUVRRT_TPE_THREADS=1024 AUVRRT_TPE_GRIDS=1 AUVRRT_TPE_STAGE_KB=8
echo "== nocost"; timeout 300 python tools/micro_run.py catalina-nocost 33554432 2>&1 | grep -v "^done" | cut -c1-110
echo "== nocost 1024 grids"; AUVRRT_TPE_THREADS=1024 AUVRRT_TPE_GRIDS=1 timeout 300 python tools/micro_run.py catalina-nocost 33554432 2>&1 | grep -v "^done" | cut -c1-110
timeout 300 python tools/micro_run.py tpt 262144 | grep -v "^done"
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'])"
