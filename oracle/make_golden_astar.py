"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/astar.npz from the UNMODIFIED reference.

Runs path_planning/astar_fixLenSOG.py:astar.astar (imported from /root/reference by oracle/harness.py,
prints swallowed) on the Catalina fixtures and stores the expansion order, the node path with its
costs, and which points smoothPath kept.  Run:  python oracle/make_golden_astar.py   (about a minute)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import harness as H  # noqa: E402

# name, start, path_len_limit, weights, velocity, number of time bins used (0 = all)
CASES = [
    ("main_like", (-212.34, 55.12), 300, (0, 10, 10, 100), 1, 0),       # main()'s weights and limit (:662, :686)
    ("short", (-212.34, 55.12), 120, (0, 10, 10, 100), 1, 0),
    ("fast_auv", (-150.0, 30.0), 90, (0, 1, 5, 50), 2, 0),
    ("west", (-300.5, 80.25), 150, (0, 10, 10, 100), 1, 0),
    ("length_only", (-100.25, 20.5), 140, (0, 10, 0, 0), 1, 0),
    ("prob_only", (-250.0, 60.0), 110, (0, 0, 0, 100), 1, 0),
    ("integer_start", (-200.0, 50.0), 160, (0, 10, 10, 100), 1, 0),
    ("near_obstacles", (-60.0, -20.0), 130, (0, 10, 10, 100), 1, 0),
    ("limit_too_small", (-212.34, 55.12), 5, (0, 10, 10, 100), 1, 0),     # smoothPath IndexError on a 1-point trajectory
    ("no_time_bin", (-212.34, 55.12), 150, (0, 10, 10, 100), 1, 2),       # findCurrSOG returns None -> AttributeError
    ("outside", (-600.0, 300.0), 100, (0, 10, 10, 100), 1, 0),            # no neighbour in bounds: returns None
]


def main():
    aref = H.load_astar_reference()
    with open(os.path.join(ROOT, "tests", "golden", "catalina_map.json")) as f:
        w = json.load(f)
    z = np.load(os.path.join(ROOT, "tests", "golden", "shark_grid.npz"))
    cells = H.lattice_cells(w["boundary"])
    data = {"names": np.array([c[0] for c in CASES]), "cells_rounded": H.round_cells(cells)}
    for name, start, limit, weights, vel, nb in CASES:
        bins = z["bins"][:nb] if nb else z["bins"]
        probs = z["probs"][:nb] if nb else z["probs"]
        t0 = time.time()
        t = H.traced_astar(aref, start, w["circles"], w["boundary"], w["habitats"], bins, cells, probs, velocity=vel,
                           path_len_limit=limit, weights=weights)
        outcome = "ok" if "nodes" in t else ("raised:" + t["raised"] if "raised" in t else "none")
        data[name + "/setup"] = np.array(list(start) + [limit] + list(weights) + [vel, nb], dtype=np.float64)
        data[name + "/outcome"] = np.array(outcome)
        data[name + "/expanded"] = t["expanded"]
        data[name + "/n_visited"] = np.array(t["n_visited"])
        data["centroid"] = t["centroid"]
        if "nodes" in t:
            for k in ("nodes", "cost_list", "smooth_keep", "smooth_path"):
                data[name + "/" + k] = t[k]
            data[name + "/cost"] = np.array(t["cost"])
        print("%-16s %-22s expanded %4d visited %4d  %.1fs" % (name, outcome, len(t["expanded"]), t["n_visited"], time.time() - t0))
    # the four-weight Cost twin (root cost.py:151-214) the A* drivers score paths with (astarAnalysis.py:37-41)
    rc = aref.mod            # astar_fixLenSOG did `from cost import Cost`: the root-level class
    M = aref.MPS
    grid = {}
    for t, (b0, b1) in enumerate(z["bins"]):
        grid[(int(b0), int(b1))] = {tuple(float(v) for v in c): float(z["probs"][t][i]) for i, c in enumerate(cells)}
    habs = [M(h[0], h[1], size=h[2]) for h in w["habitats"]]
    twin = []
    for name, start, limit, weights, vel, nb in CASES:
        if (name + "/nodes") not in data:
            continue
        nodes = data[name + "/nodes"]
        for k, (stamps, T, wts) in enumerate([(nodes[:, 3], float(nodes[-1, 3]) or 1.0, (1.0, -3.0, -3.0, -4.0)),
                                              (nodes[:, 3] * 2.5, 321.5, (0.5, 10.0, 10.0, 100.0)),     # stamps past the last bin
                                              (nodes[:, 3], -7.0, (2.0, 1.0, 0.25, 3.0)),
                                              (nodes[:, 3] * 2.5 - 40.0, 100.0, (1.0, 1.0, 1.0, 1.0))]):  # first stamp in no bin
            path = [M(r[0], r[1], traj_time_stamp=float(ts)) for r, ts in zip(nodes, stamps)]
            try:
                res = rc.Cost().habitat_shark_cost_func(path, float(nodes[-1, 2]), 1234.5, T, habs, grid, list(wts))
                row = [res[0]] + list(res[1]) + [0.0]
            except UnboundLocalError:
                row = [0.0, 0.0, 0.0, 0.0, 0.0, 1.0]
            twin.append((name, k, T, wts, row))
    data["twin/case"] = np.array([t[0] for t in twin])
    data["twin/variant"] = np.array([t[1] for t in twin])
    data["twin/T"] = np.array([t[2] for t in twin])
    data["twin/weights"] = np.array([t[3] for t in twin])
    data["twin/result"] = np.array([t[4] for t in twin])
    print("Cost twin rows:", len(twin), "raised:", int(sum(t[4][5] for t in twin)))
    out = os.path.join(ROOT, "tests", "golden", "astar.npz")
    np.savez_compressed(out, **data)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
