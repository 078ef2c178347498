#!/bin/bash
# round 2, GPU call I (2 GPUs): full GPU tests (new pins / replanning / divergence tests), config 5 at N=1 and N=2
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/i_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/i_pytest.log
grep -v "^$" gpurun_out/i_pytest.log | tail -40
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/i_bench_n1.json 2> gpurun_out/i_bench_n1.err; echo "bench n1 rc $?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/i_bench_n2.json 2> gpurun_out/i_bench_n2.err; echo "bench n2 rc $?"
python - <<'PY'
import json
for f in ("gpurun_out/i_bench_n1.json", "gpurun_out/i_bench_n2.json"):
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        print(f, d["ms_per_step"], d["value"], json.dumps(d.get("config5"))[:900])
    except Exception as ex:
        print(f, "ERR", ex)
PY
