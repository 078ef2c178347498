#!/bin/bash
# bench the planner for each lane-group size and print one summary line each
for g in 32 16 8; do
  python bench.py --steps 3 --warmup 2 --no-extras --group $g > /tmp/b_$g.json 2>/tmp/b_$g.err || tail -3 /tmp/b_$g.err
  python - "$g" <<'PY'
import json, sys
g = sys.argv[1]
try:
    d = json.loads(open('/tmp/b_%s.json' % g).read().strip().splitlines()[-1])
    print("group", g, "edges/s %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "ok", d["queries_ok"],
          "frac %.4f" % d["roofline"]["frac"], "e2e %.4g" % d["e2e"]["value"])
except Exception as e:
    print("group", g, "failed", e)
PY
done
python - <<'PY'
import sys, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/auv-sim_b200')
import numpy as np
import bench
from auvrrt import api
world, bins, probs = bench.load_world()
env = api.Env.from_map(world, bins, probs)
starts, seeds = bench.make_queries(0, 4096)
r = api.plan_batch(env, starts, seeds, api.plan_params(2048), "f32")
rec = r["records"]
print("status hist", np.bincount(rec["status"]), "depth max", rec["depth"].max(), "nodes mean", rec["n_nodes"].mean())
bad = np.where(rec["status"] != 0)[0][:5]
print(rec[bad])
PY
