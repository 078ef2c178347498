"""Drop-in for gym_rrt/envs/rrt_dubins.py of auv-sim: class Planner_RRT with the reference's
constructor and methods (/root/reference/gym_rrt/envs/rrt_dubins.py:30-503), backed by the vectorised
CUDA planner of libauvrrt.so (csrc/gym.cu) through the C ABI.  No CPU fallback.

What changes for a caller
  * random numbers come from the episode's counter-based sample sequence selected by `seed`
    (keyword-only; default drawn from Python's `random`, so `random.seed()` makes runs reproducible);
    random.choice(seq) is seq[int(u * len(seq))].
  * `replicas` (keyword-only): `planning` grows that many independent trees at once and returns the
    first one (in replica order) that reached the goal -- the batch axis the GPU is for.
  * `precision = "f64"` selects the verification build (the reference's arithmetic), "f32" the fast one.
  * an out-of-range heading no longer blocks on input() (:137-152); the last subsection is used.
"""
import math
import os
import random
import sys
import time

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from auvrrt import gym as _gym  # noqa: E402
from gym_rrt.envs.motion_plan_state_rrt import Motion_plan_state  # noqa: E402
from gym_rrt.envs.grid_cell_rrt import Grid_cell_RRT  # noqa: E402

_STATUS_EXC = {
    2: lambda: ZeroDivisionError("float division by zero"),
    3: lambda: IndexError("list index out of range"),            # random.choice([]) / env_grid[row][col]
    5: lambda: OverflowError("more nodes than node_cap"),
}


class Planner_RRT:
    """RRT planning toward a goal, one node per generate_one_node call."""

    def __init__(self, start, goal, boundary, obstacles, habitats, exp_rate=1, dist_to_end=2, diff_max=0.5, freq=50,
                 cell_side_length=2, subsections_in_cell=8, *, seed=None, replicas=1, node_cap=None, precision="f32",
                 device=0, track_counts=True):
        self.start = start
        self.goal = goal
        self.boundary_point = boundary
        self.cell_side_length = cell_side_length
        self.subsections_in_cell = subsections_in_cell
        self.obstacle_list = obstacles
        self.habitats = habitats
        self.last_path = []
        self.exp_rate = exp_rate
        self.dist_to_end = dist_to_end
        self.diff_max = diff_max
        self.freq = freq
        self.t_start = time.time()
        self.precision = precision
        self.replicas = int(replicas)
        self.seed = random.getrandbits(63) if seed is None else int(seed)
        self.node_cap = int(node_cap) if node_cap else 257
        bl, tr = boundary[0], boundary[1]
        circles = [(float(o.x), float(o.y), float(o.size)) for o in obstacles]
        self._batch = _gym.GymBatch((bl.x, bl.y, tr.x, tr.y), circles, self.replicas, exp_rate=exp_rate,
                                    dist_to_end=dist_to_end, diff_max=diff_max, freq=freq,
                                    cell_side_length=cell_side_length, subsections_in_cell=subsections_in_cell,
                                    node_cap=self.node_cap, track_counts=track_counts,
                                    precision=_gym.F64 if precision == "f64" else _gym.F32, device=device)
        self.discretize_env(cell_side_length, subsections_in_cell)
        starts = np.tile([float(start.x), float(start.y), float(start.theta)], (self.replicas, 1))
        goals = np.tile([float(goal.x), float(goal.y)], (self.replicas, 1))
        seeds = (np.arange(self.replicas, dtype=np.uint64) + np.uint64(self.seed)) & np.uint64(0x7FFFFFFFFFFFFFFF)
        self._batch.reset(starts, goals, seeds)
        self._episode = 0            # the replica `mps_list`, `env_grid` and generate_one_node talk about
        self._tree_cache = None
        self._nodes = [self.start]   # Motion_plan_state per node, extended step by step (None: rebuild from the device tree)
        self._steps = 0

    # ------------------------------------------------------------------ grid (read-only views)
    def discretize_env(self, cell_side_length, subsections_in_cell):
        """env_grid[row][col] with the reference's cell corners (:77-93)"""
        bl = self.boundary_point[0]
        self.env_grid = []
        for row in range(self._batch.rows):
            self.env_grid.append([])
            for col in range(self._batch.cols):
                base = (row * self._batch.cols + col) * subsections_in_cell
                self.env_grid[row].append(Grid_cell_RRT(bl.x + col * cell_side_length, bl.y + row * cell_side_length,
                                                        side_length=cell_side_length,
                                                        num_of_subsections=subsections_in_cell, planner=self,
                                                        flat_base=base))

    def _tree(self):
        if self._tree_cache is None:
            t = self._batch.tree(self._episode)
            nodes = self._nodes if (self._nodes is not None and len(self._nodes) == len(t["nodes"])) else None
            if nodes is None:
                nodes = [self.start]
                for i in range(1, len(t["nodes"])):
                    x, y, th, tt = t["nodes"][i]
                    m = Motion_plan_state(x, y, theta=th, traj_time_stamp=tt)
                    m.parent = nodes[t["parents"][i]]
                    nodes.append(m)
                self._nodes = nodes
            self._tree_cache = (t, nodes)
        return self._tree_cache

    @property
    def mps_list(self):
        """the tree's nodes in creation order.  generate_one_node extends the list from the step's record (the new
        node and its parent index), so an RL episode does not download the whole tree on every step."""
        if self._nodes is not None and self._tree_cache is None:
            return self._nodes
        return self._tree()[1]

    @property
    def occupied_grid_cells_array(self):
        """(row, col, subsection) in first-occupancy order"""
        ns, cols = self.subsections_in_cell, self._batch.cols
        return [(int(c) // ns // cols, int(c) // ns % cols, int(c) % ns) for c in self._tree()[0]["occupied"]]

    def _nodes_in(self, flat_id):
        t, nodes = self._tree()
        return [nodes[i] for i in np.flatnonzero(t["cells"] == flat_id)]

    def node_counts(self):
        """len(node_array) of every subsection, env_grid order (RRTEnv.convert_rrt_grid_to_1D_num_of_nodes_only)"""
        return self._batch.counts(self._episode, 1)[0]

    def print_env_grid(self):
        for row in self.env_grid:
            for grid_cell in row:
                print(grid_cell)
            print("----")

    # ------------------------------------------------------------------ planning
    def _raise(self, rec):
        if rec["status"] in _STATUS_EXC:
            raise _STATUS_EXC[int(rec["status"])]()

    def _path_of(self, q):
        rows = self._batch.path(q, cap=max(8192, 64 * self.node_cap))
        return [Motion_plan_state(x, y, theta=th) for x, y, th in rows]

    def planning(self, max_step=200, min_length=250, plan_time=True):
        """-> (path, step, seconds).  path = generate_final_course's list (goal arc end -> start) or, when
        no replica reached the goal, the last step's node / None as in the reference (:157-196)."""
        t0 = time.time()
        if max_step + self._steps + 1 > self.node_cap:
            raise OverflowError("max_step %d needs node_cap >= %d" % (max_step, max_step + self._steps + 1))
        recs = self._batch.plan(max_step)
        self._tree_cache = None
        self._nodes = None
        done = np.flatnonzero(recs["done"])
        q = int(done[0]) if len(done) else 0
        self._episode = q
        rec = recs[q]
        self._raise(rec)
        self._steps = int(rec["steps"])
        if rec["done"]:
            path = self._path_of(q)
            path[0].length = float(rec["arc_length"])
        elif rec["last_accepted"]:
            path = self.mps_list[-1]
        else:
            path = None
        return path, int(rec["steps"]), time.time() - t0

    def generate_one_node(self, grid_cell, step_num=None, min_length=250):
        """One RRTEnv.step worth of work on the sub-cell `grid_cell` (an env_grid subsection, or its flat id)
        -> (done, path | new_node | None) as the reference (:198-238)."""
        flat = int(getattr(grid_cell, "flat_id", grid_cell))
        recs = self._batch.step(np.full(self.replicas, flat, np.int32))
        self._tree_cache = None
        rec = recs[self._episode]
        self._raise(rec)
        self._steps = int(rec["steps"])
        if rec["last_parent"] < 0:
            return False, None                              # empty cell (:209-215)
        if rec["done"]:
            path = self._path_of(self._episode)
            path[0].length = float(rec["arc_length"])
            return True, path
        if rec["last_accepted"]:
            if self._nodes is not None and len(self._nodes) == int(rec["n_nodes"]) - 1:
                x, y, th, tt = (float(v) for v in rec["cand"])
                node = Motion_plan_state(x, y, theta=th, traj_time_stamp=tt)
                node.parent = self._nodes[int(rec["last_parent"])]
                self._nodes.append(node)
            else:
                self._nodes = None
                node = self.mps_list[-1]
            node.rl_state_id = step_num
            return False, node
        return False, None

    # ------------------------------------------------------------------ small host-side helpers of the class
    def angle_wrap(self, ang):
        while not (-math.pi <= ang <= math.pi):
            ang += (-2 * math.pi) if ang > math.pi else (2 * math.pi)
        return ang

    def check_within_boundary(self, mps):
        bl, tr = self.boundary_point[0], self.boundary_point[1]
        return (bl.x <= mps.x <= tr.x) and (bl.y <= mps.y <= tr.y)

    def get_distance_angle(self, start_mps, end_mps):
        dx = end_mps.x - start_mps.x
        dy = end_mps.y - start_mps.y
        return math.sqrt(dx ** 2 + dy ** 2), math.atan2(dy, dx)

    def check_collision_obstacle(self, mps, obstacleList):
        return all(self.get_distance_angle(o, mps)[0] > o.size for o in obstacleList)

    def cal_length(self, path):
        return sum(math.sqrt((path[i].x - path[i - 1].x) ** 2 + (path[i].y - path[i - 1].y) ** 2)
                   for i in range(1, len(path)))
