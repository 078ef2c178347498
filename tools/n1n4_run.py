#!/usr/bin/env python
"""driver for profiling the gym planner / lattice A* kernels: python tools/n1n4_run.py [gym|astar]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "auv-sim_b200"))
what = sys.argv[1] if len(sys.argv) > 1 else "gym"
if what == "gym":
    from auvrrt import gym
    OBST = [(12.0, 38.0, 4.0), (17.0, 34.0, 5.0), (20.0, 29.0, 4.0), (25.0, 25.0, 3.0), (29.0, 20.0, 4.0), (34.0, 17.0, 3.0), (37.0, 8.0, 5.0)]
    Q = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
    rs = np.random.default_rng(5)
    b = gym.GymBatch((0, 0, 50, 50), OBST, Q, freq=10.0, node_cap=201, precision=gym.F32)
    s = np.column_stack([rs.uniform(5, 15, Q), rs.uniform(5, 15, Q), rs.uniform(-np.pi, np.pi, Q)])
    g = np.column_stack([rs.uniform(35, 45, Q), rs.uniform(35, 45, Q)])
    for _ in range(2):
        b.reset(s, g, np.arange(Q))
        r = b.plan(200)
    print("gym", Q, r["steps"].sum(), r["done"].mean())
else:
    from auvrrt import astar
    w = json.load(open(os.path.join(ROOT, "tests/golden/catalina_map.json")))
    z = np.load(os.path.join(ROOT, "tests/golden/shark_grid.npz"))
    ga = np.load(os.path.join(ROOT, "tests/golden/astar.npz"))
    env = astar.AstarEnv(w["circles"], w["boundary"], w["habitats"], z["bins"], ga["cells_rounded"], z["probs"], centroid=ga["centroid"], cells_are_rounded=True)
    Q = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 128
    rs = np.random.default_rng(9)
    q = astar.make_queries(np.round(np.column_stack([rs.uniform(-300, -100, Q), rs.uniform(20, 90, Q)]), 2), 300.0)
    for _ in range(2):
        r = astar.astar_batch(env, q, node_cap=2048, want_paths=False)
    print("astar", Q, r["records"]["n_expanded"].sum())
