"""GPU: the drop-in Python surface (path_planning/rrt_dubins.py, cost.py) against golden vectors of
the reference.  Reads like the reference's own tests would: seed `random`, call the method."""
import json
import os
import random
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PP = os.path.join(ROOT, "auv-sim_b200", "path_planning")
from oracle import orc  # noqa: E402


@pytest.fixture(scope="module")
def world(catalina_map, shark_grid):
    saved = list(sys.path)
    sys.path.insert(0, PP)
    for n in ("rrt_dubins", "cost", "catalina", "motion_plan_state", "_world"):
        sys.modules.pop(n, None)
    import rrt_dubins, cost  # noqa
    from motion_plan_state import Motion_plan_state as M
    bins, probs = shark_grid

    class Cell:
        def __init__(self, b):
            self.bounds = tuple(b)
    cells = [Cell(c) for c in catalina_map["cells"]]
    grid = {(int(b[0]), int(b[1])): {cells[i].bounds: float(p) for i, p in enumerate(probs[t])} for t, b in enumerate(bins)}
    obstacles = [M(c[0], c[1], size=c[2]) for c in catalina_map["circles"]]
    habitats = [M(h[0], h[1], size=h[2]) for h in catalina_map["habitats"]]
    boundary = [tuple(p) for p in catalina_map["boundary"]]
    rrt = rrt_dubins.RRT(boundary, obstacles, grid, cells)
    yield {"rrt": rrt, "M": M, "cost": cost, "habitats": habitats, "grid": grid, "obstacles": obstacles,
           "oworld": orc.OracleWorld.from_map(catalina_map, bins, probs)}
    sys.path[:] = saved


def test_steer_and_collision_under_python_seeds(world, golden_dir):
    """cases i % 5 == 4 of the golden set were recorded from random.Random(i): seeding `random` the
    same way must reproduce the reference's edge through the drop-in method."""
    z = np.load(os.path.join(golden_dir, "steer_arc.npz"))
    rrt, M = world["rrt"], world["M"]
    n = 0
    for i in range(4, 400, 5):
        random.seed(i)
        p = z["parents"][i]
        new = rrt.steer(M(p[0], p[1], theta=p[2], traj_time_stamp=p[3], length=p[4]), 2, 0.5, 30, 0.5, z["velocity"][i], True)
        assert len(new.path) == z["nwp"][i]
        got = [new.x, new.y, new.theta, new.traj_time_stamp, new.length]
        assert np.allclose(got, z["leaf"][i], rtol=1e-9, atol=1e-9)
        ref = z["wp"][z["woff"][i]:z["woff"][i + 1]]
        for w, r in zip(new.path[1:], ref):
            assert np.allclose([w.x, w.y, w.theta, w.v, w.traj_time_stamp, w.length], r, rtol=1e-9, atol=1e-9)
        assert new.path[0].x == p[0]
        assert rrt.check_collision(new, rrt.obstacle_list) == bool(z["safe_fwd"][i])
        assert rrt.check_collision(new, list(reversed(rrt.obstacle_list))) == bool(z["safe_rev"][i])
        n += 1
    assert n == 80 and rrt.check_collision(None, rrt.obstacle_list) is False


def test_cost_entry_point_golden(world, golden_dir):
    z = np.load(os.path.join(golden_dir, "cost.npz"))
    M, cost = world["M"], world["cost"]
    for i in range(len(z["t_total"])):
        pts = z["pts"][z["off"][i]:z["off"][i + 1]]
        path = [M(p[0], p[1], traj_time_stamp=p[2]) for p in pts]
        w = [float(x) for x in z["weights"][i]]
        out = cost.habitat_shark_cost_func(path, float(z["t_total"][i]), world["habitats"][:int(z["n_hab"][i])], world["grid"], w)
        assert [out[0]] + out[1] == list(z["out"][i])                      # bit-exact
    with open(os.path.join(golden_dir, "cost_hand.json")) as f:
        for c in json.load(f):
            grid = {tuple(g[0]): {tuple(cb): p for cb, p in g[1]} for g in c["grid"]}
            out = cost.habitat_shark_cost_func([M(p[0], p[1], traj_time_stamp=p[2]) for p in c["points"]], c["t_total"],
                                               [M(h[0], h[1], size=h[2]) for h in c["habitats"]], grid, c["weights"])
            assert [out[0]] + out[1] == [float(v) for v in c["out"]], c


def test_get_closest_mps(world, golden_dir):
    z = np.load(os.path.join(golden_dir, "nn.npz"))
    rrt, M = world["rrt"], world["M"]
    tree = [M(a, b) for a, b in z["tree3"]]
    for q, want in zip(z["q3"], z["idx3"]):
        assert rrt.get_closest_mps(M(q[0], q[1]), tree) is tree[want]


def test_exploring_returns_reference_shaped_valid_plan(world):
    rrt, M = world["rrt"], world["M"]
    for prec in ("f64", "f32"):
        rrt.precision = prec
        res = rrt.exploring(M(-200, 0), world["habitats"], 0.5, 5, 2, 50, traj_time_stamp=True, max_plan_time=10,
                            max_traj_time=500, plan_time=True, weights=[-3, -3, -4], iterations=512, seed=3)
        assert set(res) == {"path length", "path", "cost"}
        path, split = res["path"]
        assert path[0].x == -200 and path[-1].traj_time_stamp >= 470
        assert list(split.keys())[0] == (0, 50) and sum(len(v) for v in split.values()) <= len(path)
        pts = np.array([[p.x, p.y, p.traj_time_stamp] for p in reversed(path)])
        want = orc.cost(pts, path[-1].traj_time_stamp, world["oworld"], [-3, -3, -4])
        tol = 1e-9 if prec == "f64" else 2e-4
        assert np.allclose([res["cost"][0]] + res["cost"][1], want, rtol=tol, atol=tol)
        assert orc.check_collision(np.array([[p.x, p.y] for p in path]), world["oworld"]) == 1
    # fp64 + same seed == the oracle's plan for that stream
    rrt.precision = "f64"
    res = rrt.exploring(M(-200, 0), world["habitats"], 0.5, 5, 2, 50, traj_time_stamp=True, max_traj_time=500,
                        weights=[-3, -3, -4], iterations=512, seed=9)
    o = orc.exploring(world["oworld"], [-200, 0, 0, 0, 0], orc.plan_params(512), seed=9)
    assert len(res["path"][0]) == o["n_path"] and abs(res["cost"][0] - o["result"][1]) < 1e-9
    # replicas: the best of several trees is at least as good as the first
    r4 = rrt.exploring(M(-200, 0), world["habitats"], 0.5, 5, 2, 50, traj_time_stamp=True, max_traj_time=500,
                       weights=[-3, -3, -4], iterations=512, seed=9, replicas=8)
    assert r4["cost"][0] <= res["cost"][0] + 1e-12
    # nearest-node mode
    rn = rrt.exploring(M(-200, 0), world["habitats"], 0.5, 5, 2, 50, traj_time_stamp=True, max_traj_time=500,
                       plan_time=False, weights=[-3, -3, -4], iterations=768, seed=1)
    assert rn["path"][0][-1].traj_time_stamp >= 470


def test_exceptions_mirror_reference(world):
    rrt, M = world["rrt"], world["M"]
    with pytest.raises(TypeError):                       # start inside an obstacle: no node ever reaches the horizon
        rrt.exploring(M(9.39, -3.2), world["habitats"], 0.5, 5, 2, 50, traj_time_stamp=True, max_traj_time=500,
                      iterations=64, seed=0)
    with pytest.raises(ValueError):
        e = M(0, 0); e.path = []
        rrt.check_collision(e, rrt.obstacle_list)


def test_replanning_runs(world):
    rrt, M = world["rrt"], world["M"]
    rrt.precision = "f32"
    random.seed(5)
    traj, time_dict, cost = rrt.replanning(M(-200, 0), list(world["habitats"]), 0.25, 120.0, 0.1, [-3, -3, -4])
    assert len(traj) > 10 and len(time_dict) >= 2 and len(cost) == 2 and len(cost[1]) == 3
    ts = [p.traj_time_stamp for p in traj]
    assert ts[-1] > 200


def test_replanning_matches_reference_trace(world, golden_dir):
    """RRT.replanning (rrt_dubins.py:51-90) against traces of the unmodified reference (tests/golden/pins.npz): every
    inner exploring call on the slot-addressed stream of seed + k with the same iteration budget.  The drop-in detaches
    `initial` from the previous tree; the reference walks on into the old tree when the start point was one of its
    NODES (has_parent in the fixture), so segments are compared up to the first such start, the whole run when none."""
    rrt, M = world["rrt"], world["M"]
    z = np.load(os.path.join(golden_dir, "pins.npz"))
    meta = json.loads(str(z["meta"]))
    rrt.precision = "f64"
    checked_all = 0
    try:
        for ci, c in enumerate(meta):
            tag = "rp%d_" % ci
            rrt.replan_seed, rrt.replan_iterations = c["seed"], c["iterations"]
            traj, time_dict, cost = rrt.replanning(M(c["start"][0], c["start"][1]), list(world["habitats"]), c["budget"], c["length"],
                                                   c["interval"], [-3, -3, -4])
            hp = z[tag + "has_parent"]
            n_ok = int(np.argmax(hp)) + 1 if hp.any() else len(hp)       # the flagged segment itself still starts alike
            off = z[tag + "first_off"]
            want = z[tag + "first"]
            got = np.array([[p.x, p.y, p.theta, p.v, p.traj_time_stamp, p.length] for p in traj])
            n_pts = int(off[n_ok - 1]) if hp.any() else int(off[-1])       # points of the segments before the flagged one
            assert len(got) >= n_pts
            assert np.allclose(got[:n_pts], want[:n_pts], rtol=1e-9, atol=1e-9), ci
            assert [len(v[1]) for v in time_dict.values()][:n_ok - 1 if hp.any() else None] == \
                list(z[tag + "habitats_left"])[:n_ok - 1 if hp.any() else None]
            if not hp.any():
                assert len(got) == len(z[tag + "traj"]) and np.allclose(got, z[tag + "traj"], rtol=1e-9, atol=1e-9)
                assert np.allclose([cost[0]] + list(cost[1]), z[tag + "cost"], rtol=1e-9, atol=1e-12)
                checked_all += 1
        assert checked_all >= 1
    finally:
        rrt.replan_seed = rrt.replan_iterations = None
        rrt.precision = "f32"
