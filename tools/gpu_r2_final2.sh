#!/bin/bash
# round 2, end of round (second pass, after the CTA-shape changes of k_plan and the slim thread-per-tree state):
# full GPU suite, smoke(), k_plan capture, the full bench line and the reference arm (1 GPU)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/final_pytest.log
tail -3 gpurun_out/final_pytest.log
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none -k regex:k_plan -s 2 -c 1 -o gpurun_out/f_plan python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/f_ncu_plan.log 2>&1
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
