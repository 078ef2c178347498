#!/bin/bash
# round 2, end of round (fourth pass, after the 896-thread shape of k_plan): smoke(), k_plan and thread-per-tree captures,
# the full bench line and the reference arm (1 GPU).  The full GPU suite ran on this build in call Z12 (74 passed).
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none -k regex:k_plan -s 2 -c 1 -o gpurun_out/f_plan python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/f_ncu_plan.log 2>&1
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none -k regex:k_plan_tpt -c 1 -o gpurun_out/f_tpt python tools/micro_run.py tpt 262144 > gpurun_out/f_ncu_tpt.log 2>&1
timeout 1500 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; tail -c 300 gpurun_out/final_bench_n1.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench_n1.json').read().strip().splitlines()[-1])
print('ms',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])
print(d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
for k,v in d.get('extras',{}).items():
    if isinstance(v,dict): print(k, {kk:vv for kk,vv in v.items() if kk in ('edges_per_s','frac','error','plans_per_s','identical_booleans','steps_per_s','queries_per_s')})
c=d.get('config5'); print({k:c.get(k) for k in ('seconds','plans_per_s','best_query')})
r=json.loads(open('gpurun_out/final_bench_ref.json').read().strip().splitlines()[-1])
print('reference arm', r['value'], r['cpu_baseline']['cores'], r['ms_per_step'])
PY
