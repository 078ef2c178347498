"""Drop-in path_planning/astar_fixLenSOG.py: signatures (CPU) and results through the drop-in against the
golden queries of the unmodified reference (GPU)."""
import inspect
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PP = os.path.join(ROOT, "auv-sim_b200", "path_planning")
_NAMES = ("astar_fixLenSOG", "rrt_dubins", "cost", "catalina", "motion_plan_state", "_world")


@pytest.fixture(scope="module")
def mods():
    saved = list(sys.path)
    saved_mods = {n: sys.modules.pop(n) for n in _NAMES if n in sys.modules}
    sys.path.insert(0, PP)
    import astar_fixLenSOG, motion_plan_state  # noqa
    yield astar_fixLenSOG, motion_plan_state.Motion_plan_state
    sys.path[:] = saved
    for n in _NAMES:
        sys.modules.pop(n, None)
    sys.modules.update(saved_mods)


def _solver(mods, world, start, velocity, n_bins=0):
    A, M = mods
    bins = world["bins"][:n_bins] if n_bins else world["bins"]
    probs = world["probs"][:n_bins] if n_bins else world["probs"]
    grid = {}
    for t, (b0, b1) in enumerate(bins):
        grid[(int(b0), int(b1))] = {tuple(float(v) for v in c): float(probs[t][i]) for i, c in enumerate(world["cells"])}
    obst = [M(c[0], c[1], size=c[2]) for c in world["circles"]]
    bnd = [M(p[0], p[1]) for p in world["boundary"]]
    hab = [M(h[0], h[1], size=h[2]) for h in world["habitats"]]
    return A.astar((float(start[0]), float(start[1])), obst, bnd, hab, grid, {}, velocity)


def test_signatures_match_reference(mods):
    A, M = mods
    sig = lambda f: [p for p in inspect.signature(f).parameters if p not in ("cell_list", "device")]
    assert sig(A.astar.__init__) == ["self", "start", "obstacleList", "boundaryList", "habitatList", "sharkGrid", "shark_dict",
                                     "AUV_velocity"]
    assert sig(A.astar.astar) == ["self", "pathLenLimit", "weights", "shark_traj"]
    n = A.Node(None, (1.0, 2.0))
    assert (n.parent, n.position, n.g, n.h, n.f, n.cost, n.pathLen, n.time_stamp) == (None, (1.0, 2.0), 0, 0, 0, 0, 0, 0)
    assert A.euclidean_dist((0, 0), (3, 4)) == 5.0
    with pytest.raises(ValueError):
        A.astar((0.0, 0.0), [], [M(0, 0), M(1, 0), M(0, 1)], [], {}, {}, 1)


def test_no_gpu_fails_loudly(mods, catalina_map, shark_grid):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from auvrrt import AuvrrtError
    world = dict(catalina_map, bins=shark_grid[0], probs=shark_grid[1])
    s = _solver(mods, world, (-212.34, 55.12), 1)
    with pytest.raises(AuvrrtError, match="no CUDA device"):
        s.astar(100, [0, 10, 10, 100], {})


@pytest.mark.gpu
def test_dropin_matches_golden(mods, astar_golden, catalina_map):
    A, M = mods
    world, cases = astar_golden
    world = dict(world, cells=catalina_map["cells"])
    # the drop-in rounds the raw cell bounds itself (Python round) and computes the centroid itself
    for c in cases:
        s = _solver(mods, world, c.start, c.velocity if c.velocity != int(c.velocity) else int(c.velocity), c.n_bins)
        w = [int(v) for v in c.weights]
        if c.outcome == "ok":
            r = s.astar(c.limit, w, {})
            assert r["path length"] == len(c.smooth_path) and r["cost"] == c.cost
            assert r["cost list"] == list(c.cost_list)
            got = np.array([[n.position[0], n.position[1], n.pathLen, n.time_stamp, n.cost, n.f] for n in r["node"]])
            assert np.array_equal(got, c.nodes)
            assert np.array_equal(np.array([[m.x, m.y, m.traj_time_stamp] for m in r["path"]]), c.smooth_path)
            assert r["node"][0].parent is None and r["node"][-1].parent is r["node"][-2]
        elif c.outcome == "none":
            assert s.astar(c.limit, w, {}) is None
        else:
            exc = {"IndexError": IndexError, "AttributeError": AttributeError, "TypeError": TypeError}[c.outcome.split(":")[1]]
            with pytest.raises(exc):
                s.astar(c.limit, w, {})
    ok = [c for c in cases if c.outcome == "ok" and c.velocity == 1 and c.n_bins == 0][:4]
    s = _solver(mods, world, ok[0].start, 1)
    many = s.astar_many([c.start for c in ok], [c.limit for c in ok], [c.weights for c in ok])
    for r, c in zip(many, ok):
        assert r["cost"] == c.cost and r["path length"] == len(c.smooth_path)


@pytest.mark.gpu
def test_cost_twin_matches_reference(mods, astar_golden, catalina_map):
    """root-level cost.py Cost.habitat_shark_cost_func (the A* drivers' scorer) through the drop-in, bit for bit"""
    A, M = mods
    import cost as dropin_cost
    world, cases = astar_golden
    z = np.load(os.path.join(ROOT, "tests", "golden", "astar.npz"))
    by_name = {c.name: c for c in cases}
    grid = {}
    for t, (b0, b1) in enumerate(world["bins"]):
        grid[(int(b0), int(b1))] = {tuple(float(v) for v in c): float(world["probs"][t][i]) for i, c in enumerate(catalina_map["cells"])}
    habs = [M(h[0], h[1], size=h[2]) for h in world["habitats"]]
    n_ok = n_raise = 0
    for name, k, T, wts, want in zip(z["twin/case"], z["twin/variant"], z["twin/T"], z["twin/weights"], z["twin/result"]):
        nodes = by_name[str(name)].nodes
        stamps = [nodes[:, 3], nodes[:, 3] * 2.5, nodes[:, 3], nodes[:, 3] * 2.5 - 40.0][int(k)]
        path = [M(r[0], r[1], traj_time_stamp=float(ts)) for r, ts in zip(nodes, stamps)]
        args = (path, float(nodes[-1, 2]), 1234.5, float(T), habs, grid, [float(v) for v in wts])
        if want[5]:
            with pytest.raises(UnboundLocalError):
                dropin_cost.Cost().habitat_shark_cost_func(*args)
            n_raise += 1
        else:
            got = dropin_cost.Cost().habitat_shark_cost_func(*args)
            assert got[0] == want[0] and list(got[1]) == list(want[1:5]), (name, k)
            n_ok += 1
    assert n_ok >= 20 and n_raise >= 4
    with pytest.raises(ZeroDivisionError):
        dropin_cost.Cost().habitat_shark_cost_func(path[:0], 1.0, 2.0, 0.0, habs, grid, [1, 1, 1, 1])
