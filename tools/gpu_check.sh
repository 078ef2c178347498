#!/bin/bash
# parity suite + planner benches (cooperative mb variants, thread-per-tree at several batch sizes)
python -m pytest tests -m gpu -q 2>&1 | tail -6
bash tools/gpu_variants.sh "$@"
for q in 4096 65536 131072; do
  python bench.py --steps 2 --warmup 1 --no-extras --group 1 --queries $q > /tmp/b.json 2>/tmp/b.err || tail -3 /tmp/b.err
  python - "$q" <<'PY'
import json, sys
try:
    d = json.loads(open('/tmp/b.json').read().strip().splitlines()[-1])
    print("tpt Q", sys.argv[1], "edges/s %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "ok", d["queries_ok"], "e2e %.4g" % d["e2e"]["value"])
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
