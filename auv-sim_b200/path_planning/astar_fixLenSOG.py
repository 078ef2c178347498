"""Drop-in for path_planning/astar_fixLenSOG.py of auv-sim: class `astar` with the reference's constructor
and `astar(pathLenLimit, weights, shark_traj)` (/root/reference/path_planning/astar_fixLenSOG.py:114-657),
backed by the batched CUDA planner of libauvrrt.so (csrc/astar.cu).  No CPU fallback.

Same results as the reference, bit for bit (the planner is deterministic fp64): the returned dict has
"path length", "path" (smoothPath's trajectory), "cost", "cost list" and "node" exactly as the reference
builds them.  What changes for a caller
  * `sharkGrid` must be passed ({(t0, t1): {cell.bounds: p}}, e.g. from rrt_dubins.createSharkGrid); the
    reference's fallback of reading 'path_planning/shark_data/AUVGrid_prob_500_straight.csv' against a
    shapely cell split (:129-130) needs `cell_list=` here (see sharkOccupancyGrid.splitCell);
  * the diagnostic print()s are gone;
  * `astar_many(queries)` (extra) plans a batch of (start, pathLenLimit, weights) in one launch;
  * exceptions mirror the reference: AstarLookupError, which is an AttributeError (time stamp outside every
    time bin, :489), a TypeError (point outside every cell, :602) and an IndexError (:419, :532, :652);
    `None` when the open list runs empty.
"""
import csv
import math

import numpy as np

import _world  # noqa: F401  (puts the package directory on sys.path)
from auvrrt import astar as _astar
from motion_plan_state import Motion_plan_state


def euclidean_dist(point1, point2):
    dx = abs(point1[0] - point2[0])
    dy = abs(point1[1] - point2[1])
    return math.sqrt(dx * dx + dy * dy)


def createSharkGrid(filepath, cell_list):
    """CSV rows {"time bin": "(t0, t1)", "grid": "[p0, p1, ...]"} -> {(t0, t1): {cell.bounds: p}} (:31-49; the
    last column is dropped, as the reference does)"""
    out = {}
    with open(filepath, newline='') as csvfile:
        for row in csv.DictReader(csvfile):
            a, b = row["time bin"].split(", ")
            key = (int(a[1:]), int(b[:-1]))
            vals = row["grid"][1:-1].split(", ")
            out[key] = {cell_list[i].bounds: float(vals[i]) for i in range(len(vals) - 1)}
    return out


class AstarLookupError(AttributeError, TypeError, IndexError):
    """The reference fails with AttributeError (no time bin holds the time stamp, :489), TypeError (no cell holds
    the point, :602) or IndexError (:419, :532, :652); the kernel reports them as one status, so this is all three."""


class Node:
    """a node of the search graph (:99-112)"""

    def __init__(self, parent=None, position=None):
        self.parent = parent
        self.position = position
        self.g = 0
        self.h = 0
        self.f = 0
        self.cost = 0
        self.pathLen = 0
        self.time_stamp = 0


class astar:
    def __init__(self, start, obstacleList, boundaryList, habitatList, sharkGrid, shark_dict, AUV_velocity, *,
                 cell_list=None, device=0):
        self.start = start
        self.velocity = AUV_velocity
        self.obstacle_list = obstacleList
        self.boundary_list = boundaryList
        self.habitat_list = habitatList
        self.visited_nodes = np.zeros([600, 600])
        self.cell_list = cell_list
        if sharkGrid == {}:
            if cell_list is None:
                raise ValueError("pass sharkGrid, or cell_list= to label 'path_planning/shark_data/"
                                 "AUVGrid_prob_500_straight.csv' (the reference splits the boundary with shapely)")
            sharkGrid = createSharkGrid('path_planning/shark_data/AUVGrid_prob_500_straight.csv', cell_list)
        self.sharkGrid = sharkGrid
        self.sharkDict = shark_dict
        self.device = device
        self._env = None

    def _environment(self):
        if self._env is None:
            bins = [[float(k[0]), float(k[1])] for k in self.sharkGrid]
            first = next(iter(self.sharkGrid.values()), {})
            cells = [list(map(float, b)) for b in first]
            probs = [[float(g[b]) for b in first] for g in self.sharkGrid.values()]
            self._env = _astar.AstarEnv([[o.x, o.y, o.size] for o in self.obstacle_list],
                                        [[c.x, c.y] for c in self.boundary_list],
                                        [[h.x, h.y, h.size] for h in self.habitat_list], bins, cells, probs, device=self.device)
        return self._env

    _EXC = {3: lambda: AstarLookupError("time stamp outside every time bin, point outside every cell, or index out of range"),
            5: lambda: OverflowError("more than 4096 lattice nodes / 512 path points")}

    def _result(self, r, i):
        rec = r["records"][i]
        if rec["status"] == 1:
            return None                                     # open list ran empty
        if rec["status"] in self._EXC:
            raise self._EXC[int(rec["status"])]()
        rows = r["paths"][i][:rec["n_path"]]
        nodes, prev = [], None
        for x, y, plen, ts, cost, f in rows:
            n = Node(prev, (x, y))
            n.pathLen, n.time_stamp, n.cost, n.g, n.f = plen, int(ts), cost, cost, f
            n.h = f - cost
            nodes.append(n)
            prev = n
        traj = [Motion_plan_state(n.position[0], n.position[1], traj_time_stamp=round(n.time_stamp, 2)) for n in nodes]
        smooth = [m for m, k in zip(traj, r["keep"][i][:rec["n_path"]]) if k]
        return {"path length": len(smooth), "path": smooth, "cost": float(rec["cost"]),
                "cost list": [n.cost for n in reversed(nodes)], "node": nodes}

    def astar(self, pathLenLimit, weights, shark_traj):
        """-> {"path length", "path", "cost", "cost list", "node"} or None (:551-657)"""
        q = _astar.make_queries([self.start], float(pathLenLimit), [float(w) for w in weights[:4]], float(self.velocity))
        r = _astar.astar_batch(self._environment(), q, path_cap=512)
        return self._result(r, 0)

    def astar_many(self, starts, pathLenLimits, weights_list):
        """extra: one launch for len(starts) independent queries -> list of result dicts / None"""
        n = len(starts)
        q = np.zeros(n, _astar.ASTAR_QUERY_DTYPE)
        q["start"] = np.asarray(starts, dtype=np.float64).reshape(n, 2)
        q["path_len_limit"] = np.broadcast_to(np.asarray(pathLenLimits, dtype=np.float64), (n,))
        q["weights"] = np.broadcast_to(np.asarray(weights_list, dtype=np.float64), (n, 4))
        q["velocity"] = float(self.velocity)
        r = _astar.astar_batch(self._environment(), q, path_cap=512)
        return [self._result(r, i) for i in range(n)]

    # small host-side helpers of the class (:143-176, :205-221)
    def euclidean_dist(self, point1, point2):
        dx = abs(point1[0] - point2[0])
        dy = abs(point1[1] - point2[1])
        return dx * dx + dy * dy

    def get_distance_angle(self, start_mps, end_mps):
        dx = end_mps.x - start_mps.x
        dy = end_mps.y - start_mps.y
        return math.sqrt(dx ** 2 + dy ** 2), math.atan2(dy, dx)

    def collision_free(self, position, obstacleList):
        return all(math.sqrt((position[0] - o.x) ** 2 + (position[1] - o.y) ** 2) > o.size for o in obstacleList)

    def with_in_time_bin(self, time_bin, curr_time_stamp):
        return time_bin[0] <= curr_time_stamp <= time_bin[1]

    def get_indices(self, x_in_meters, y_in_meters):
        return (int(x_in_meters + 500), int(y_in_meters + 200))
