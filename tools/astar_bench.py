"""Quick throughput look at the batched lattice A* (not the bench contract; see bench.py extras)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "auv-sim_b200")); sys.path.insert(0, ROOT)
from auvrrt import astar
from oracle import orc

w = json.load(open(os.path.join(ROOT, "tests/golden/catalina_map.json")))
z = np.load(os.path.join(ROOT, "tests/golden/shark_grid.npz"))
g = np.load(os.path.join(ROOT, "tests/golden/astar.npz"))
env = astar.AstarEnv(w["circles"], w["boundary"], w["habitats"], z["bins"], g["cells_rounded"], z["probs"], centroid=g["centroid"], cells_are_rounded=True)
rs = np.random.default_rng(1)
for Q in (1, 148 * 4, 148 * 32, 148 * 128):
    q = astar.make_queries(np.round(np.column_stack([rs.uniform(-300, -100, Q), rs.uniform(20, 90, Q)]), 2), 300.0)
    astar.astar_batch(env, q, want_paths=False)
    t0 = time.perf_counter(); r = astar.astar_batch(env, q, want_paths=False); dt = time.perf_counter() - t0
    rec = r["records"]
    print(f"Q={Q}: {dt*1e3:.2f} ms, {Q/dt:.3e} queries/s, expansions/s {rec['n_expanded'].sum()/dt:.3e}, ok {np.mean(rec['status']==0):.2f}, mean expanded {rec['n_expanded'].mean():.0f}", flush=True)
ow = orc.astar_world(w["circles"], w["boundary"], g["centroid"], w["habitats"], z["bins"], g["cells_rounded"], z["probs"])
Qc = 64
q = astar.make_queries(np.round(np.column_stack([rs.uniform(-300, -100, Qc), rs.uniform(20, 90, Qc)]), 2), 300.0)
oq = np.column_stack([q["start"], q["path_len_limit"], q["weights"], q["velocity"]])
t0 = time.perf_counter(); recs, cost, st = orc.astar_batch(ow, oq); dt = time.perf_counter() - t0
print(f"oracle port, {orc.num_threads()} threads: {Qc/dt:.3e} queries/s, expansions/s {recs[:,0].sum()/dt:.3e}")
