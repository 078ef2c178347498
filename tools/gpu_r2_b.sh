#!/bin/bash
# round 2, GPU call B: GPU tests, thread-per-edge kernel after the branch-light lookup rework, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/b_pytest.log
tail -15 gpurun_out/b_pytest.log
for v in catalina catalina-nocost catalina-allpairs; do
  timeout 300 python tools/micro_run.py $v 33554432 >> gpurun_out/b_micro.log 2>&1
done
cat gpurun_out/b_micro.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc $?"; cat gpurun_out/b_bench.json | cut -c1-400
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_edges_arc_tpe -s 1 -c 1 -o gpurun_out/b_tpe python tools/micro_run.py catalina 8388608 > gpurun_out/b_ncu_tpe.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_plan -s 2 -c 1 -o gpurun_out/b_plan python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/b_ncu_plan.log 2>&1
ls -la gpurun_out | tail -8
