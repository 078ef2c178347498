#!/bin/bash
# round 2, GPU call Z13: thread-per-edge kernel, edges per thread per batch 8 (default) / 16 / 4
for v in default ept16 ept4; do echo "== $v"; if [ $v = default ]; then timeout 300 python tools/micro_run.py catalina 33554432 2>&1 | grep -v "^done" | cut -c1-100; else AUVRRT_LIB=$PWD/gpurun_variants/libauvrrt_$v.so timeout 300 python tools/micro_run.py catalina 33554432 2>&1 | grep -v "^done" | cut -c1-100; fi; done
