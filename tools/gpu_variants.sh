#!/bin/bash
# bench the planner with each prebuilt library variant (AUVRRT_LIB) and print one line each
for lib in "$@"; do
  AUVRRT_LIB=$PWD/auv-sim_b200/auvrrt/$lib python bench.py --steps 3 --warmup 2 --no-extras > /tmp/b.json 2>/tmp/b.err || tail -3 /tmp/b.err
  python - "$lib" <<'PY'
import json, sys
try:
    d = json.loads(open('/tmp/b.json').read().strip().splitlines()[-1])
    print(sys.argv[1], "edges/s %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "ok", d["queries_ok"],
          "frac %.4f" % d["roofline"]["frac"], "e2e %.4g" % d["e2e"]["value"])
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
