#!/bin/bash
# round 2, GPU call U: smoke(); k_plan with / without the straight-line single-candidate tests in the warp-per-edge evaluation; captures
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2
b() { timeout 600 python bench.py --steps 10 --warmup 3 --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench ms/step %.3f value %.4g e2e %.4g' % (d['ms_per_step'],d['value'],d['e2e']['value']))"; }
for i in 1 2; do
echo "== default"; b
for v in e0 h0 eh0; do echo "== $v"; AUVRRT_LIB=$PWD/gpurun_variants/libauvrrt_$v.so b; done
done
timeout 600 ncu --set full --clock-control none -k regex:k_plan -s 2 -c 1 -o gpurun_out/u_plan python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/u_ncu_plan.log 2>&1
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none -k regex:k_edges_arc_tpe -s 1 -c 1 -o gpurun_out/u_tpe python tools/micro_run.py catalina 8388608 > gpurun_out/u_ncu_tpe.log 2>&1
ls -la gpurun_out/*.ncu-rep
for sh in 256 512 1024; do for g in 0 1; do echo "== config4 nocost threads $sh grids $g"; AUVRRT_TPE_THREADS=$sh AUVRRT_TPE_GRIDS=$g timeout 300 python tools/micro_run.py config4-nocost 33554432 2>&1 | grep -v "^done" | cut -c1-100; done; done
for sh in 256 1024; do echo "== config4 cost threads $sh"; AUVRRT_TPE_THREADS=$sh timeout 300 python tools/micro_run.py config4 33554432 2>&1 | grep -v "^done" | cut -c1-100; done
