#!/bin/bash
# round 2, GPU call C: thread-per-edge kernel with out-of-line slow paths / bin cursor / FASTENV
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plan_variants.py -m gpu -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c_pytest.log
tail -12 gpurun_out/c_pytest.log
for v in catalina catalina-nocost catalina-allpairs; do
  timeout 300 python tools/micro_run.py $v 33554432 >> gpurun_out/c_micro.log 2>&1
  AUVRRT_TPE_MINB=4 timeout 300 python tools/micro_run.py $v 33554432 >> gpurun_out/c_micro.log 2>&1
done
AUVRRT_TPE_STAGE_KB=24 timeout 300 python tools/micro_run.py catalina 33554432 >> gpurun_out/c_micro.log 2>&1
AUVRRT_TPE_STAGE_KB=24 AUVRRT_TPE_MINB=4 timeout 300 python tools/micro_run.py catalina 33554432 >> gpurun_out/c_micro.log 2>&1
grep -v "^done" gpurun_out/c_micro.log
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_edges_arc_tpe -s 1 -c 1 -o gpurun_out/c_tpe python tools/micro_run.py catalina 8388608 > gpurun_out/c_ncu_tpe.log 2>&1
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_edges_arc_tpe -s 1 -c 1 -o gpurun_out/c_tpe_ap python tools/micro_run.py catalina-allpairs 8388608 > gpurun_out/c_ncu_tpe_ap.log 2>&1
ls -la gpurun_out | tail -6
