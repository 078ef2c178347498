#!/bin/bash
# round 2, GPU call W: end-of-round evidence on one GPU: launch list, captures of the remaining shipped kernels, the full bench line
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/w_launches.csv python bench.py --steps 2 --warmup 1 --micro-edges 16777216 --config5-queries 65536 > gpurun_out/w_launches_bench.log 2>&1
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none -k regex:k_edges_arc_tpe -s 1 -c 1 -o gpurun_out/w_tpe_ap python tools/micro_run.py catalina-allpairs 8388608 > gpurun_out/w_ncu_tpe_ap.log 2>&1
MICRO_REPS=2 timeout 600 ncu --set full --clock-control none -k regex:k_edges_arc_tpe -s 1 -c 1 -o gpurun_out/w_tpe_ap4 python tools/micro_run.py config4-allpairs 2097152 > gpurun_out/w_ncu_tpe_ap4.log 2>&1
AUVRRT_EDGES_BRUTE=1 timeout 600 ncu --section SpeedOfLight --section ComputeWorkloadAnalysis --section WarpStateStats --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis --section SchedulerStats --section InstructionStats --clock-control none -k regex:k_edges_dubins -s 1 -c 1 -o gpurun_out/w_dubins python tools/micro_run.py edges 16777216 > gpurun_out/w_ncu_dubins.log 2>&1
timeout 600 ncu --section SpeedOfLight --section ComputeWorkloadAnalysis --section WarpStateStats --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis --section SchedulerStats --section InstructionStats --clock-control none -k regex:k_nn_partial -s 1 -c 1 -o gpurun_out/w_nn python tools/micro_run.py nn 134217728 > gpurun_out/w_ncu_nn.log 2>&1
timeout 600 ncu --section SpeedOfLight --section ComputeWorkloadAnalysis --section WarpStateStats --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis --section SchedulerStats --section InstructionStats --clock-control none -k regex:k_plan -s 1 -c 1 -o gpurun_out/w_mode1 python tools/micro_run.py mode1 4096 > gpurun_out/w_ncu_mode1.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 1500 python bench.py > gpurun_out/w_bench_n1.json 2> gpurun_out/w_bench_n1.err; tail -c 300 gpurun_out/w_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/w_bench_n1.json').read().strip().splitlines()[-1])
print('ms',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])
print(d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['sample'][:120])
for k,v in d.get('extras',{}).items():
    if isinstance(v,dict): print(k, {kk:vv for kk,vv in v.items() if kk in ('edges_per_s','frac','error','plans_per_s','identical_booleans','equivalent_frac','steps_per_s','queries_per_s')})
print(d.get('config5'))
PY
