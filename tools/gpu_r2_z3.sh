#!/bin/bash
# round 2, GPU call Z3: thread-per-tree planner, CTA size 128 (default) / 256 / 512
mkdir -p gpurun_out
echo "== t128 (default)"; timeout 300 python tools/micro_run.py tpt 262144 | grep -v "^done"
for v in t256 t512; do echo "== $v"; AUVRRT_LIB=$PWD/gpurun_variants/libauvrrt_$v.so timeout 300 python tools/micro_run.py tpt 262144 | grep -v "^done"; done
for v in t256 t512; do echo "== $v 524288"; AUVRRT_LIB=$PWD/gpurun_variants/libauvrrt_$v.so timeout 300 python tools/micro_run.py tpt 524288 | grep -v "^done"; done
echo "== t128 524288"; timeout 300 python tools/micro_run.py tpt 524288 | grep -v "^done"
