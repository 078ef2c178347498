#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_occupancy.py tests/test_dropin_gym.py tests/test_gpu_dropin.py -m gpu -q > gpurun_out/j_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/j_pytest.log
tail -5 gpurun_out/j_pytest.log
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none -k regex:k_edges_arc_tpe -s 1 -c 1 -o gpurun_out/j_tpe_ap python tools/micro_run.py catalina-allpairs 8388608 > gpurun_out/j_ncu_tpe_ap.log 2>&1
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none -k regex:k_edges_arc_tpe -s 1 -c 1 -o gpurun_out/j_tpe python tools/micro_run.py catalina 8388608 > gpurun_out/j_ncu_tpe.log 2>&1
ls -la gpurun_out
