#!/bin/bash
# round 2, last call: the new many-circles parity test, the sanitizers on the extended driver
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "allpairs or shapes or edges" > gpurun_out/last_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/last_pytest.log
tail -3 gpurun_out/last_pytest.log
bash tools/gpu_sanitize.sh 2>/dev/null | grep -v "^+" | tail -12
