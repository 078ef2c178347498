"""Drop-in for path_planning/cost.py of auv-sim.

`habitat_shark_cost_func` (the planner's path cost, /root/reference/path_planning/cost.py:145-207)
and `habitat_shark_cost_point` (:209-241) keep their call signatures and run on the GPU through the
C ABI (auvrrt_cost / auvrrt_cost_point, fp64 verification build: same result as the reference,
including its `mps.x <= cell_bound[3]` typo, the skipped no-bin waypoints and CPython's
compensated sum()).  There is no CPU fallback for these two.  The four small helpers the planner
never calls are plain host loops with the reference's signatures.
"""
import math

import numpy as np

import _world
from _world import api

_envs = _world.EnvCache()


def _env_for(habitats, shark_dict, fingerprint=None):
    habs = _world.circles_of(habitats)
    fp = fingerprint if fingerprint is not None else _world.grid_fingerprint(shark_dict)
    return _envs.get(np.zeros((0, 3)), np.zeros((0, 2)), habs, fp, lambda: _world.grid_of(shark_dict))


def habitat_shark_cost_func(path, total_traj_time, habitats, shark_dict, weight):
    """cost = [w1 * visited/len(habitats) + w2 * time-in-habitat/T + w3 * shark prob/T, [c0, c1, c2]]"""
    env = _env_for(habitats, shark_dict)
    pts = np.array([[p.x, p.y, p.traj_time_stamp] for p in path], dtype=np.float64).reshape(-1, 3)
    out = api.cost(env, [pts], [float(total_traj_time)], [float(w) for w in weight[:3]], precision="f64")[0]
    return [float(out[0]), [float(out[1]), float(out[2]), float(out[3])]]


def habitat_shark_cost_point(mps, habitats, visited, AUVGrid, weight):
    """single-state variant; `visited` is returned unchanged, as in the reference (:234 compares
    instead of assigning)"""
    env = _env_for(habitats, {(0, 0): AUVGrid} if AUVGrid else {})      # keyed by the grid's CONTENT (_world.grid_fingerprint)
    vis = [1 if visited[i] else 0 for i in range(len(habitats))]
    s = api.cost_point(env, [[mps.x, mps.y]], vis, 0, [float(w) for w in weight[:3]], precision="f64")[0]
    return float(s), visited


# ---- helpers outside the planner's hot path (host loops, reference signatures) -------------------
def _dist(a, b):
    return math.sqrt((a.x - b.x) ** 2 + (a.y - b.y) ** 2)


def test_cost_func(path, length, bonus_area, weights=[1, 0]):
    cost = weights[0] * length
    for mps in path:
        if _dist(bonus_area[0], mps) <= bonus_area[0].size:
            cost -= weights[1]
    return cost


def habitat_num_cost_func(path, length, habitats, weights=[1, 1]):
    count = sum(1 for h in habitats if any(_dist(h, mps) <= h.size for mps in path))
    return weights[0] * length - weights[1] * count


def cost_of_edge(new_node, habitat_open_list, habitat_closed_list, weights):
    def inside(habs):
        return int(any(math.sqrt((new_node.position[0] - h.x) ** 2 + (new_node.position[1] - h.y) ** 2) <= h.size
                       for h in habs))
    d_2 = inside(habitat_open_list + habitat_closed_list)
    d_3 = inside(habitat_closed_list)
    return [-weights[1] * d_2 - weights[2] * d_3, d_2, d_3]


def habitat_time_cost_func(path, length, habitats, dist, weights=[1, -1, -1]):
    cost = [0 for _ in weights]
    cost[0] = weights[0] * length / dist
    visited = [False] * len(habitats)
    d = dist
    for i, h in enumerate(habitats):
        for mps in path:
            d = _dist(h, mps)                 # the reference reuses `dist` as its loop variable
            if d <= h.size:
                visited[i] = True
                cost[2] += weights[2]
    cost[2] = cost[2] / (0.5 * d)
    cost[1] = weights[1] * sum(visited) / len(habitats)
    return [sum(cost), cost]


class Cost:
    """Drop-in for the root-level cost.py class of auv-sim (/root/reference/cost.py:10-214), the four-weight twin
    the A* drivers use (`from cost import Cost`, astar_fixLenSOG.py:16, astarAnalysis.py:37-41).  The small
    methods are the host helpers above; `habitat_shark_cost_func` runs on the GPU (auvrrt_cost, fp64
    verification build) and reproduces the twin's differences from path_planning/cost.py: the extra
    w1 * length / peri term, no guard for a time stamp outside every bin (the previous waypoint's bin is
    reused; UnboundLocalError when there is none yet, :187-191) and the unconditional division by
    total_traj_time (:203-204)."""

    def __init__(self):
        self.cost = 0

    def test_cost_func(self, path, length, bonus_area, weights=[1, 0]):
        return test_cost_func(path, length, bonus_area, weights)

    def habitat_num_cost_func(self, path, length, habitats, weights=[1, 1]):
        return habitat_num_cost_func(path, length, habitats, weights)

    def cost_of_edge(self, new_node, habitat_open_list, habitat_closed_list, weights):
        return cost_of_edge(new_node, habitat_open_list, habitat_closed_list, weights)

    def habitat_time_cost_func(self, path, length, habitats, dist, weights=[1, -1, -1]):
        return habitat_time_cost_func(path, length, habitats, dist, weights)

    def habitat_shark_cost_func(self, path, length, peri, total_traj_time, habitats, shark_dict, weight):
        """-> [sum(cost), [w1 * length / peri, w2 * visited / len(habitats), w3-term / T, w4-term / T]]"""
        w1, w2, w3, w4 = weight[0], weight[1], weight[2], weight[3]
        cost = [0 for _ in range(len(weight))]
        cost[0] = w1 * length / peri
        bins = list(shark_dict)
        pts, last = [], None
        for mps in path:
            t = mps.traj_time_stamp
            if any(t >= b[0] and t <= b[1] for b in bins):
                last = t
            elif last is None:
                raise UnboundLocalError("cannot access local variable 'temp_time' where it is not associated with a value")
            pts.append([mps.x, mps.y, last])          # a stamp outside every bin reuses the previous waypoint's bin
        if total_traj_time == 0:
            raise ZeroDivisionError("float division by zero" if isinstance(total_traj_time, float) else "division by zero")
        env = _env_for(habitats, shark_dict)
        pts = np.array(pts, dtype=np.float64).reshape(-1, 3)
        out = api.cost(env, [pts], [float(total_traj_time)], [float(w2), float(w3), float(w4)], precision="f64")[0]
        c1, c2, c3 = float(out[1]), float(out[2]), float(out[3])
        if total_traj_time < 0:                       # the kernel only normalises for T > 0 (path_planning/cost.py:193)
            c2, c3 = c2 / total_traj_time, c3 / total_traj_time
        cost[1], cost[2], cost[3] = c1, c2, c3
        return [sum(cost), cost]
