#!/bin/bash
# round 2, GPU call Z: thread-per-tree planner with the slim per-tree shared state: 1 slot per thread (8 / 10 CTAs per SM), 2 slots (6 / 5 CTAs)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -q -x -k "tpt or thread_per_tree or multi or sharded" > gpurun_out/z_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/z_pytest.log
tail -3 gpurun_out/z_pytest.log
echo "== spt1 minb8 (default)"; timeout 300 python tools/micro_run.py tpt 262144 | grep -v "^done"
for v in spt1m10 spt2 spt2m5; do echo "== $v"; AUVRRT_LIB=$PWD/gpurun_variants/libauvrrt_$v.so timeout 300 python tools/micro_run.py tpt 262144 | grep -v "^done"; done
echo "== spt2 524288"; AUVRRT_LIB=$PWD/gpurun_variants/libauvrrt_spt2.so timeout 300 python tools/micro_run.py tpt 524288 | grep -v "^done"
echo "== spt1 524288"; timeout 300 python tools/micro_run.py tpt 524288 | grep -v "^done"
AUVRRT_LIB=$PWD/gpurun_variants/libauvrrt_spt2.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tpt or thread_per_tree" 2>&1 | tail -2
