"""CPU-only: the oracle (oracle/auvrrt_oracle.c) and the host-side helpers against tests/golden/pins.npz, outputs of the
UNMODIFIED reference (oracle/make_golden_pins.py): check_collision_obstacle, habitat_shark_cost_point,
habitat_shark_cost_func on a filtered shark dict, and splitPath (via the split counts in exploring.npz)."""
import os
import sys

import numpy as np
import pytest

from oracle import orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def pins():
    return np.load(os.path.join(GOLDEN, "pins.npz"))


@pytest.fixture(scope="module")
def oworld(catalina_map, shark_grid):
    return orc.OracleWorld.from_map(catalina_map, shark_grid[0], shark_grid[1])


def test_check_collision_obstacle_matches_reference(pins, oworld):
    """rrt_dubins.py:551-556: per-obstacle test with the obstacle's own radius (no running minimum), <= collides"""
    got = np.array([orc.check_collision_obstacle(p[0], p[1], oworld) for p in pins["k2_points"]], dtype=np.uint8)
    assert np.array_equal(got, pins["k2_safe"]) and 0 < pins["k2_safe"].sum() < len(got)


def test_cost_point_matches_reference(pins, oworld):
    """cost.py:209-241; `visited[i] == True` there is a comparison, not an assignment: visited comes back unchanged"""
    assert np.array_equal(pins["c2_visited_after"], pins["c2_visited"])
    for i, p in enumerate(pins["c2_points"]):
        got = orc.cost_point(p[0], p[1], oworld, pins["c2_visited"][i], int(pins["c2_tb"][i]), pins["c2_weights"][i])
        assert got == pins["c2_out"][i], i
    assert (pins["c2_out"] != 0).sum() > 100


def test_cost_with_bin_filter_matches_reference(pins, oworld):
    """the planner's bin filter (rrt_dubins.py:161-166) hands cost.habitat_shark_cost_func a sub-dict: a waypoint whose
    time falls in a dropped bin takes the next kept bin that contains it, or is skipped (cost.py:173-179)"""
    off = pins["bm_off"]
    for i in range(len(off) - 1):
        p = pins["bm_pts"][off[i]:off[i + 1]]
        got = orc.cost(p, float(pins["bm_T"][i]), oworld, [-3, -3, -4], bin_mask=pins["bm_mask"][i])
        assert np.array_equal(got, pins["bm_out"][i]), i


def test_split_path_matches_reference(exploring_golden):
    """RRT.splitPath (rrt_dubins.py:590-602) of the drop-in class on the golden optimal paths: the bucket sizes the
    reference produced (result["path"][1])"""
    z, meta = exploring_golden
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pp = os.path.join(root, "auv-sim_b200", "path_planning")
    saved = list(sys.path)
    sys.path.insert(0, pp)
    for n in ("rrt_dubins", "cost", "catalina", "motion_plan_state", "_world"):
        sys.modules.pop(n, None)
    try:
        import rrt_dubins
        from motion_plan_state import Motion_plan_state
        n = 0
        for m in meta:
            tag = m["tag"]
            if tag + "_split_counts" not in z:
                continue
            path = [Motion_plan_state(r[0], r[1], theta=r[2], v=r[3], traj_time_stamp=r[4], length=r[5]) for r in z[tag + "_path"]]
            split = rrt_dubins.RRT.splitPath(None, path, 50, [0.0, 500.0])
            assert [len(v) for v in split.values()] == list(z[tag + "_split_counts"]), tag
            assert list(split.keys()) == [(50.0 * i, 50.0 * (i + 1)) for i in range(10)]
            n += 1
        assert n >= 20
    finally:
        sys.path[:] = saved
        for n_ in ("rrt_dubins", "cost", "catalina", "motion_plan_state", "_world"):
            sys.modules.pop(n_, None)
