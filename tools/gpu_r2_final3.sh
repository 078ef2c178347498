#!/bin/bash
# round 2, end of round (third pass, after the packed time-bin rows of the thread-per-tree planner): full GPU suite, smoke(),
# thread-per-tree capture, the full bench line (1 GPU)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/final_pytest.log
tail -3 gpurun_out/final_pytest.log
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -1
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none -k regex:k_plan_tpt -c 1 -o gpurun_out/f_tpt python tools/micro_run.py tpt 262144 > gpurun_out/f_ncu_tpt.log 2>&1
timeout 1500 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; tail -c 300 gpurun_out/final_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench_n1.json').read().strip().splitlines()[-1])
print('ms',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])
for k,v in d.get('extras',{}).items():
    if isinstance(v,dict): print(k, {kk:vv for kk,vv in v.items() if kk in ('edges_per_s','frac','error','plans_per_s','identical_booleans','steps_per_s','queries_per_s')})
c=d.get('config5'); print({k:c.get(k) for k in ('seconds','plans_per_s','best_query')})
PY
