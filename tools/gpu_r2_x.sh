#!/bin/bash
# round 2, GPU call X (gpurun --gpus 8): multi-GPU bench lines (headline weak scaling + config 5 strong scaling) at N = 2, 4, 8
mkdir -p gpurun_out
NG=${1:-8}
for N in 2 4 8; do
  if [ $N -le $NG ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/x_bench_n$N.out 2> gpurun_out/x_bench_n$N.err
    grep '^{' gpurun_out/x_bench_n$N.out | tail -1 > gpurun_out/x_bench_n$N.json
    python - <<PY
import json
d=json.loads(open('gpurun_out/x_bench_n$N.json').read())
print('N', d['n_gpus'], 'ms', round(d['ms_per_step'],3), 'value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'])
c=d.get('config5') or {}
print('  config5', {k:c.get(k) for k in ('queries','seconds','plans_per_s','best_query','sample_records_equal_single_gpu','error')})
PY
  fi
done
