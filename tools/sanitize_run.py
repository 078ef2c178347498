#!/usr/bin/env python
"""small workload over every kernel for compute-sanitizer (memcheck / racecheck / synccheck)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "auv-sim_b200"))
from auvrrt import api  # noqa

world = json.load(open(os.path.join(ROOT, "tests", "golden", "catalina_map.json")))
g = np.load(os.path.join(ROOT, "tests", "golden", "shark_grid.npz"))
env = api.Env.from_map(world, g["bins"], g["probs"])
Q = 48
starts = np.tile([-200.0, 0.0, 0.0, 0.0, 0.0], (Q, 1)); seeds = np.arange(Q)
for prec in ("f32", "f64"):
    for G in (32, 16, 8, 1):
        for mode in (0, 1, 2):
            pp = api.plan_params(96, mode=mode, group=G, trace=True, path_cap=512 if G != 1 else 0, freq=40.0 if mode == 0 else 30.0)
            r = api.plan_batch(env, starts, seeds, pp, prec)
            assert (r["records"]["status"] <= 1).all()
    rs = np.random.RandomState(0)
    parents = np.stack([rs.uniform(-300, -100, 200), rs.uniform(-60, 100, 200), rs.uniform(-6, 6, 200), rs.uniform(0, 400, 200), rs.uniform(0, 500, 200)], 1)
    api.edges_arc(env, parents, np.arange(200), [2.0, 0.5, 30.0, 0.5, 2.0], prec)
    api.edges_arc_cost(env, parents, np.arange(200), [2.0, 0.5, 30.0, 0.5, 2.0], -4.0, prec)
    # round 2: the thread-per-edge kernel in its other CTA shapes (grid plane in shared memory), all pairs, warp per edge
    for shape, grids in (("512", "0"), ("1024", "1"), ("1024", "0")):
        os.environ["AUVRRT_TPE_THREADS"] = shape; os.environ["AUVRRT_TPE_GRIDS"] = grids
        api.edges_arc_cost(env, parents, np.arange(200), [2.0, 0.5, 30.0, 0.5, 2.0], -4.0, prec)
        api.edges_arc(env, parents, np.arange(200), [2.0, 0.5, 30.0, 0.5, 2.0], prec)
    os.environ.pop("AUVRRT_TPE_THREADS"); os.environ.pop("AUVRRT_TPE_GRIDS")
    os.environ["AUVRRT_EDGES_BRUTE"] = "1"
    api.edges_arc_cost(env, parents, np.arange(200), [2.0, 0.5, 30.0, 0.5, 2.0], -4.0, prec)
    os.environ["AUVRRT_EDGES_BRUTE"] = "0"
    os.environ["AUVRRT_EDGES_VARIANT"] = "warp"
    api.edges_arc_cost(env, parents, np.arange(200), [2.0, 0.5, 30.0, 0.5, 2.0], -4.0, prec)
    os.environ.pop("AUVRRT_EDGES_VARIANT")
    # round 2: Dubins-RRT with best parent (mode 3), Dubins edges with cost on
    p3 = api.plan_params(48, mode=3, v=1.0, max_traj_time=120.0, dubins_rho=3.0, dubins_eta=8.0, near_radius=15.0, dubins_w=6,
                         trace=True, path_cap=512)
    r3 = api.plan_batch(env, starts[:8], seeds[:8], p3, prec)
    assert (r3["records"]["status"] <= 1).all()
    q0 = np.stack([rs.uniform(-400, 50, 300), rs.uniform(-100, 120, 300), rs.uniform(-3, 3, 300)], 1)
    q1 = q0 + rs.uniform(-20, 20, (300, 3))
    api.edges_dubins(env, q0, q1, 1.0, 20, prec)
    api.edges_dubins_cost(env, q0, q1, 1.0, 20, 1.0, -4.0, prec)
    os.environ["AUVRRT_EDGES_BRUTE"] = "1"
    api.edges_dubins(env, q0, q1, 1.0, 20, prec)
    os.environ["AUVRRT_EDGES_BRUTE"] = "0"
    api.steer_dubins(q0, q1, 1.0, 9, prec)
    api.nn(rs.uniform(-400, 50, (5001, 2)), rs.uniform(-400, 50, (7, 2)), prec)
    api.collide(env, [parents[:5, :2], parents[:0, :2], parents[:70, :2]], prec)
    api.collide_points(env, parents[:50, :2], prec)
    api.cost(env, [np.concatenate([parents[:40, :2], parents[:40, 3:4]], 1)], [100.0], [-3, -3, -4], precision=prec)
    api.cost_point(env, parents[:9, :2], [0] * 10, 1, [-3, -3, -4], prec)
# round 2: the all-pairs arc variant on a dense world (circle quads, two waypoints per pass, the direct-formula fallback)
rs = np.random.RandomState(7)
dense = np.stack([rs.uniform(-467.4, 82.4, 300), rs.uniform(-153.5, 191.2, 300), rs.uniform(1, 5, 300)], 1)
env_d = api.Env(circles=dense, boundary=world["boundary"], habitats=world["habitats"], bins=g["bins"], cells=world["cells"], probs=g["probs"])
par_d = np.stack([rs.uniform(-467, 82, 3000), rs.uniform(-153, 191, 3000), rs.uniform(-3, 3, 3000), rs.uniform(0, 400, 3000), np.zeros(3000)], 1)
os.environ["AUVRRT_EDGES_BRUTE"] = "1"
for prec in ("f32", "f64"):
    api.edges_arc_cost(env_d, par_d, np.arange(3000), [2.0, 0.5, 30.0, 0.5, 2.0], -4.0, prec)
os.environ["AUVRRT_EDGES_BRUTE"] = "0"
api.edges_arc_cost(env_d, par_d, np.arange(3000), [2.0, 0.5, 30.0, 0.5, 2.0], -4.0, "f32")
env_d.close()
z = np.load(os.path.join(ROOT, "tests", "golden", "occupancy.npz"))
polys = [z["cell_xy"][z["cell_off"][i]:z["cell_off"][i + 1]] for i in range(len(z["cell_off"]) - 1)]
toff, trk = z["c1_toff"], z["c1_trk"]
api.occupancy_grid(polys, z["bounds"], 10.0, 20.0, 25.0, [trk[toff[i]:toff[i + 1]] for i in range(len(toff) - 1)])
# gym planner (both builds; stepping, one-launch planning, agent actions, counts, path)
from auvrrt import gym  # noqa: E402
from auvrrt import astar  # noqa: E402
OBST = [(12.0, 38.0, 4.0), (17.0, 34.0, 5.0), (20.0, 29.0, 4.0), (25.0, 25.0, 3.0), (29.0, 20.0, 4.0)]
for prec in (gym.F32, gym.F64):
    Qg = 70
    rs = np.random.default_rng(1)
    b = gym.GymBatch((0, 0, 50, 50), OBST, Qg, freq=10.0, node_cap=61, track_counts=True, precision=prec)
    b.reset(np.column_stack([rs.uniform(5, 15, Qg), rs.uniform(5, 15, Qg), rs.uniform(-3, 3, Qg)]),
            np.column_stack([rs.uniform(35, 45, Qg), rs.uniform(35, 45, Qg)]), np.arange(Qg))
    for it in range(5):
        cnt = b.counts()
        b.step(np.argmax(cnt > 0, axis=1).astype(np.int32), full_candidates=bool(it & 1))
    r = b.plan(55)
    for q in np.flatnonzero(r["done"])[:3]:
        b.path(int(q)); b.tree(int(q))
    b.close()
# lattice A*
ga = np.load(os.path.join(ROOT, "tests", "golden", "astar.npz"))
aenv = astar.AstarEnv(world["circles"], world["boundary"], world["habitats"], g["bins"], ga["cells_rounded"], g["probs"],
                      centroid=ga["centroid"], cells_are_rounded=True)
qa = astar.make_queries([[-212.34, 55.12], [-150.0, 30.0], [-600.0, 300.0], [-300.5, 80.25], [-60.0, -20.0]], 90.0)
ra = astar.astar_batch(aenv, qa, trace=True)
assert (ra["records"]["status"][[0, 1, 3, 4]] == 0).all()
astar.astar_batch(aenv, qa, node_cap=40, want_paths=False)
aenv.close()
print("sanitize workload done, launches:", api.launch_count())
