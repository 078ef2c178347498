"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/gym_plan.npz from the UNMODIFIED reference.

Runs gym_rrt/envs/rrt_dubins.py:Planner_RRT (imported from /root/reference by oracle/harness.py)
on the pre-generated sample sequence and stores what every step decided.  Needs /root/reference;
the fixture it writes travels with the repo.  Run:  python oracle/make_golden_gym.py
"""
import math
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import harness as H  # noqa: E402

KEYS = ["parent", "nwp", "accepted", "done", "n_nodes", "n_occupied", "n_uniforms", "cand", "occupied", "nodes",
        "path"]


def episodes():
    """(name, start, goal, boundary, obstacles, kwargs, action_seed or None)"""
    out = []
    for i in range(20):                                   # Planner_RRT.planning, main()'s world (:505-519)
        rs = random.Random(1000 + i)
        start = (rs.uniform(5, 15), rs.uniform(5, 15), rs.uniform(-math.pi, math.pi))
        goal = (rs.uniform(35, 45), rs.uniform(35, 45))
        kw = dict(freq=[50, 10, 30][i % 3], cell_side_length=[2, 2, 5, 1][i % 4], subsections_in_cell=[8, 8, 4, 6][i % 4],
                  max_step=[200, 120][i % 2])
        out.append(("plan%02d" % i, start, goal, (0.0, 0.0, 50.0, 50.0), H.GYM_MAIN_OBSTACLES, kw, None))
    out.append(("plan_main", (10.0, 10.0, 0.0), (35.0, 40.0), (0.0, 0.0, 50.0, 50.0), H.GYM_MAIN_OBSTACLES,
                dict(freq=50, max_step=200), None))                                    # main()'s own query (:506-507)
    out.append(("plan_free", (10.0, 10.0, 0.3), (40.0, 30.0), (0.0, 0.0, 50.0, 50.0), [], dict(freq=10, max_step=50), None))
    for i in range(2):                                    # boundary not at the origin: list[-k] indexing (:115-154)
        rs = random.Random(2000 + i)
        start = (rs.uniform(-15, -5), rs.uniform(-15, -5), rs.uniform(-1, 2))
        out.append(("plan_neg%d" % i, start, (20.0, 22.0), (-20.0, -20.0, 30.0, 30.0),
                    [(0.0, 5.0, 4.0), (8.0, -2.0, 3.0), (12.0, 15.0, 5.0)], dict(freq=20, max_step=150), None))
    for i in range(6):                                    # RRTEnv.step: the agent picks the cell (rrt_env.py:206-224)
        rs = random.Random(3000 + i)
        start = (rs.uniform(5, 15), rs.uniform(5, 15), rs.uniform(-math.pi, math.pi))
        goal = (rs.uniform(35, 45), rs.uniform(35, 45))
        kw = dict(freq=10, cell_side_length=[2, 5][i % 2], subsections_in_cell=[8, 4][i % 2], max_step=150)
        out.append(("step%02d" % i, start, goal, (0.0, 0.0, 50.0, 50.0), H.GYM_MAIN_OBSTACLES, kw, 77 + i))
    return out


def main():
    g = H.load_gym_reference()
    data = {}
    names = []
    for seed, (name, start, goal, boundary, obstacles, kw, aseed) in enumerate(episodes()):
        rng = H.StreamPlayer(seed=seed)
        actions = None
        if aseed is not None:
            ar = random.Random(aseed)

            def actions(pl, ar=ar):
                if ar.random() < 0.15:     # sometimes an arbitrary (most likely empty) cell
                    return (ar.randrange(len(pl.env_grid)), ar.randrange(len(pl.env_grid[0])),
                            ar.randrange(pl.subsections_in_cell))
                return ar.choice(pl.occupied_grid_cells_array)
        t = H.traced_gym_planning(g, start, goal, boundary, obstacles, rng=rng, actions=actions, **kw)
        names.append(name)
        data[name + "/setup"] = np.array(list(start) + list(goal) + list(boundary) +
                                         [kw.get("freq", 50), kw.get("cell_side_length", 2),
                                          kw.get("subsections_in_cell", 8), kw["max_step"], seed, t["steps"],
                                          int(t["found"]), t["goal_arc_length"]], dtype=np.float64)
        data[name + "/obstacles"] = np.array(obstacles, dtype=np.float64).reshape(-1, 3)
        data[name + "/grid_shape"] = t["grid_shape"]
        nz = np.flatnonzero(t["counts"].reshape(-1))
        data[name + "/counts_nz"] = np.stack([nz, t["counts"].reshape(-1)[nz]], 1).astype(np.int32)
        if aseed is not None:
            data[name + "/actions"] = t["actions"]
        for k in KEYS:
            data[name + "/" + k] = t[k]
        print(name, "steps", t["steps"], "found", t["found"], "nodes", len(t["nodes"]), "path", len(t["path"]))
    data["names"] = np.array(names)
    out = os.path.join(ROOT, "tests", "golden", "gym_plan.npz")
    np.savez_compressed(out, **data)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
