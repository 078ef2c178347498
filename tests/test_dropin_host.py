"""CPU: host-side logic of the drop-in modules (no GPU compute): boundary types, map conversion,
CSV format, path bookkeeping, and that the planner fails loudly without a GPU."""
import json
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PP = os.path.join(ROOT, "auv-sim_b200", "path_planning")


@pytest.fixture(scope="module")
def mods():
    saved = list(sys.path)
    sys.path.insert(0, PP)
    for n in ("rrt_dubins", "cost", "catalina", "motion_plan_state", "_world"):
        sys.modules.pop(n, None)
    import rrt_dubins, cost, catalina, motion_plan_state  # noqa
    yield rrt_dubins, cost, catalina, motion_plan_state
    sys.path[:] = saved
    for n in ("rrt_dubins", "cost", "catalina", "motion_plan_state", "_world"):
        sys.modules.pop(n, None)


def test_signatures_match_reference(mods):
    import inspect
    rrt_dubins, cost, catalina, mps = mods
    # keyword-only extras of the drop-in (iteration budget, seed, replicas, the Dubins steer of planner mode 3) do not
    # change how the reference's callers call it
    extras = ("iterations", "seed", "replicas", "steer", "dubins_rho", "dubins_eta", "near_radius", "dubins_w")
    assert all(inspect.signature(rrt_dubins.RRT.exploring).parameters[k].kind is inspect.Parameter.KEYWORD_ONLY for k in extras)
    sig = lambda f: [p for p in inspect.signature(f).parameters if p not in extras]
    R = rrt_dubins.RRT
    assert sig(R.__init__) == ["self", "boundary", "obstacles", "sharkGrid", "cell_list", "exp_rate", "dist_to_end", "diff_max", "freq"]
    assert sig(R.exploring) == ["self", "initial", "habitats", "plot_interval", "bin_interval", "v", "shark_interval",
                                "traj_time_stamp", "max_plan_time", "max_traj_time", "plan_time", "weights"]
    assert sig(R.replanning) == ["self", "start", "habitats", "plan_time_budget", "traj_time_length", "replan_time_interval", "weight"]
    assert sig(R.steer) == ["self", "mps", "dist_to_end", "diff_max", "freq", "min_dist", "velocity", "traj_time_stamp"]
    assert sig(R.check_collision) == ["self", "mps", "obstacleList"]
    assert sig(R.get_closest_mps) == ["self", "ran_mps", "mps_list"]
    assert sig(cost.habitat_shark_cost_func) == ["path", "total_traj_time", "habitats", "shark_dict", "weight"]
    assert sig(cost.habitat_shark_cost_point) == ["mps", "habitats", "visited", "AUVGrid", "weight"]
    d = inspect.signature(R.__init__).parameters
    assert (d["exp_rate"].default, d["dist_to_end"].default, d["diff_max"].default, d["freq"].default) == (1, 2, 0.5, 30)
    m = mps.Motion_plan_state(1, 2)
    assert (m.z, m.theta, m.v, m.w, m.traj_time_stamp, m.plan_time_stamp, m.size, m.parent, m.path, m.length) == \
        (0, 0, 0, 0, 0, 0, 0, None, [], 0)


def test_catalina_map_matches_reference_fixture(mods, catalina_map):
    _, _, catalina, _ = mods
    env = catalina.create_environs(catalina.OBSTACLES, catalina.BOUNDARIES, catalina.BOATS, catalina.HABITATS)
    got = np.array([[o.x, o.y, o.size] for o in env[0] + env[2]])
    assert np.abs(got - np.array(catalina_map["circles"])).max() < 1e-9       # metres; two Vincenty restatements
    assert np.abs(np.array([[b.x, b.y] for b in env[1]]) - np.array(catalina_map["boundary"])).max() < 1e-9
    assert np.abs(np.array([[h.x, h.y, h.size] for h in env[3]]) - np.array(catalina_map["habitats"])).max() < 1e-9
    assert len(env[0]) == 11 and len(env[2]) == 16 and len(env[3]) == 10 and len(env[1]) == 5


def test_create_shark_grid_format(mods, tmp_path):
    rrt_dubins = mods[0]

    class Cell:
        def __init__(self, b):
            self.bounds = b
    cells = [Cell((0.0, 0.0, 10.0, 10.0)), Cell((0.0, 10.0, 10.0, 20.0)), Cell((10.0, 0.0, 20.0, 10.0))]
    p = tmp_path / "g.csv"
    p.write_text('time bin,grid\n"(0, 50)","[0.5, 0.25, 0.125]"\n"(50, 100)","[0.0, 1.0, 2.0]"\n')
    g = rrt_dubins.createSharkGrid(str(p), cells)
    assert list(g.keys()) == [(0, 50), (50, 100)]
    assert g[(0, 50)] == {(0.0, 0.0, 10.0, 10.0): 0.5, (0.0, 10.0, 10.0, 20.0): 0.25, (10.0, 0.0, 20.0, 10.0): 0.125}
    with pytest.raises(IndexError):                       # row longer than cell_list (reference :628)
        rrt_dubins.createSharkGrid(str(p), cells[:2])
    import _world
    bins, cc, probs = _world.grid_of(g)
    assert bins.tolist() == [[0, 50], [50, 100]] and cc.shape == (3, 4) and probs[1].tolist() == [0.0, 1.0, 2.0]


def test_path_bookkeeping(mods):
    rrt_dubins, _, _, mps = mods
    M = mps.Motion_plan_state
    r = rrt_dubins.RRT(None, [], {}, [])
    pts = [M(0, 0, traj_time_stamp=t) for t in (0, 10, 50, 50.5, 120, 149.9, 150, 400)]
    sp = r.splitPath(pts, 50, [0, 160])
    assert list(sp.keys()) == [(0, 50), (50, 100), (100, 150)]
    assert [len(v) for v in sp.values()] == [3, 1, 3]            # first matching interval wins at t = 50, 150
    habs = [M(0, 0, size=1), M(10, 0, size=1), M(0.5, 0, size=1)]
    left = r.removeHabitat(habs, [M(0.2, 0), M(10, 0.5)])
    assert left is habs and [(h.x, h.y) for h in left] == [(0.5, 0)]
    a, b, c = M(0, 0), M(1, 0), M(2, 0)
    b.parent, c.parent = a, b
    b.path, c.path = [a, M(0.5, 0)], [b, M(1.5, 0)]
    assert [p.x for p in r.generate_final_course(c)] == [2, 1.5, 1, 0.5, 0]
    assert r.get_distance_angle(M(0, 0), M(3, 4)) == (5.0, math.atan2(4, 3))
    with pytest.raises(AttributeError):
        r.planning()


def test_no_gpu_fails_loudly(mods):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    rrt_dubins, cost, _, mps = mods
    M = mps.Motion_plan_state
    r = rrt_dubins.RRT([(0, 0), (10, 0), (10, 10), (0, 10)], [M(5, 5, size=1)], {}, [])
    from auvrrt import AuvrrtError
    with pytest.raises(AuvrrtError, match="no CUDA device"):
        r.check_collision(M(1, 1), r.obstacle_list) if False else r.exploring(
            M(1, 1), [], 0.5, 5, 2, 50, traj_time_stamp=True, max_traj_time=100)
    with pytest.raises(AuvrrtError, match="no CUDA device"):
        cost.habitat_shark_cost_func([M(1, 1)], 1.0, [M(0, 0, size=2)], {}, [-1, -1, -1])


def test_env_cache_is_keyed_by_grid_content(mods):
    """the world-model cache must notice a NEW grid with the same time bins and cells but other probabilities (CPython
    reuses the id() of a freed dict) and an in-place mutation"""
    import _world
    cells = [(0.0, 0.0, 10.0, 10.0), (10.0, 0.0, 20.0, 10.0)]
    fps = set()
    for i in range(12):
        g = {(0, 100): {c: 0.1 * (i + 1) + j for j, c in enumerate(cells)}}       # fresh dict every time, same keys
        fps.add(_world.grid_fingerprint(g))
    assert len(fps) == 12
    g = {(0, 100): {c: 0.5 for c in cells}}
    a = _world.grid_fingerprint(g)
    assert a == _world.grid_fingerprint({(0, 100): {c: 0.5 for c in cells}})      # same content, other object: same key
    g[(0, 100)][cells[0]] = 0.75
    assert _world.grid_fingerprint(g) != a                                        # in-place mutation changes the key
    assert _world.grid_fingerprint({}) is None


def test_sparse_bins_are_rejected_with_directions(mods):
    """per-bin dicts that list different cells (what SharkOccupancyGrid.convert emits when a cell's probability is 0 in
    some bin) cannot share one device cell index: the error says what to do"""
    import _world
    g = {(0, 50): {(0.0, 0.0, 10.0, 10.0): 0.5, (10.0, 0.0, 20.0, 10.0): 0.25}, (50, 100): {(10.0, 0.0, 20.0, 10.0): 0.125}}
    with pytest.raises(NotImplementedError, match="keep_zero_cells=True"):
        _world.grid_of(g)
