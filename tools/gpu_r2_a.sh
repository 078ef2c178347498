#!/bin/bash
# round 2, GPU call A: full GPU test suite on the new sample sequence + first numbers of the thread-per-edge kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
for v in catalina catalina-nocost catalina-allpairs; do
  timeout 300 python tools/micro_run.py $v 33554432 >> gpurun_out/a_micro.log 2>&1
done
AUVRRT_TPE_STAGE_KB=24 timeout 300 python tools/micro_run.py catalina 33554432 >> gpurun_out/a_micro.log 2>&1
AUVRRT_TPE_STAGE_KB=100 timeout 300 python tools/micro_run.py catalina 33554432 >> gpurun_out/a_micro.log 2>&1
AUVRRT_EDGES_VARIANT=warp timeout 300 python tools/micro_run.py catalina 33554432 >> gpurun_out/a_micro.log 2>&1
cat gpurun_out/a_micro.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc $?"
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_edges_arc_tpe -s 1 -c 1 -o gpurun_out/a_tpe python tools/micro_run.py catalina 8388608 > gpurun_out/a_ncu_tpe.log 2>&1
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_edges_arc_tpe -s 1 -c 1 -o gpurun_out/a_tpe_allpairs python tools/micro_run.py catalina-allpairs 8388608 > gpurun_out/a_ncu_tpe_ap.log 2>&1
ls -la gpurun_out | tail -12
