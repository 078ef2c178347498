#!/bin/bash
# round 2, GPU call S: k_plan look-ahead prefetch on/off; tpt prefetch; Dubins cost-on edges; full bench line with the new extras
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plan_variants.py tests/test_gpu_dropin.py -m gpu -q -x > gpurun_out/s_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/s_pytest.log
tail -3 gpurun_out/s_pytest.log
for i in 1 2; do
echo "== prefetch on"; timeout 600 python bench.py --steps 10 --warmup 3 --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'])"
echo "== prefetch off"; AUVRRT_LIB=$PWD/gpurun_variants/libauvrrt_nopf.so timeout 600 python bench.py --steps 10 --warmup 3 --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'])"
done
timeout 300 python tools/micro_run.py tpt 262144 | grep -v "^done"
timeout 1500 python bench.py --steps 20 --warmup 3 --config5-queries 262144 > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; tail -c 600 gpurun_out/s_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s_bench.json').read().strip().splitlines()[-1])
print('ms',d['ms_per_step'],'value',d['value'])
for k,v in d.get('extras',{}).items():
    if isinstance(v,dict): print(k, {kk:vv for kk,vv in v.items() if kk in ('edges_per_s','frac','error','plans_per_s','roofline','identical_booleans','equivalent_frac')})
PY
