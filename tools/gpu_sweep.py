#!/usr/bin/env python
"""run bench.py under several configurations and print one summary line each.
usage: gpu_sweep.py "label;ENV=V ENV2=V;--args ..." ..."""
import json
import os
import subprocess
import sys

for spec in sys.argv[1:]:
    label, envs, args = (spec.split(";") + ["", ""])[:3]
    env = dict(os.environ)
    for kv in envs.split():
        k, v = kv.split("=", 1)
        if k == "AUVRRT_LIB" and not v.startswith("/"):
            v = os.path.join(os.getcwd(), "auv-sim_b200", "auvrrt", v)
        env[k] = v
    r = subprocess.run([sys.executable, "bench.py", "--no-extras"] + args.split(), env=env, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print("%-28s edges/s %.4g  ms/step %.2f  ok %d  frac %.4f  e2e %.4g" % (
            label, d["value"], d["ms_per_step"], d["queries_ok"], d["roofline"]["frac"], d["e2e"]["value"]), flush=True)
    except Exception as e:
        print(label, "FAILED", e, r.stderr[-400:], flush=True)
