"""Drop-in for path_planning/rrt_dubins.py of auv-sim: class RRT with the reference's constructor
and method signatures (/root/reference/path_planning/rrt_dubins.py:22-630), backed by the CUDA
kernels of libauvrrt.so through the C ABI.  There is no CPU fallback: without the library or a GPU
every planner call raises.

What changes for a caller
  * `exploring` replaces the reference's wall-clock budget (`max_plan_time` seconds of a Python
    loop, :107-118) by an iteration budget: `iterations` steer calls (default
    max_plan_time x 410, the rate the reference reaches with the cost active)
    on a counter-based uniform stream selected by `seed` (default: drawn from `random`, so
    `random.seed()` makes runs reproducible).  `replicas` > 1 grows that many independent trees in
    one batch and returns the best plan.  These three are keyword-only extras; everything else is
    the reference's signature and return shape.
  * the single-edge twins (`steer`, `check_collision`, `get_closest_mps`, ...) draw from Python's
    `random` in exactly the reference's order and evaluate on the GPU in the fp64 verification
    build, so under the same `random.seed()` they return what the reference returns.
  * exceptions mirror the reference: TypeError when no node reached the horizon (:174),
    ZeroDivisionError (:270, :281), KeyError (:123-127).
"""
import csv
import math
import random

import numpy as np

import _world
from _world import api
from motion_plan_state import Motion_plan_state
from cost import habitat_shark_cost_func

# steer calls per second of max_plan_time: what the pure-Python reference manages with the shark
# grid active (measured)
REFERENCE_STEER_CALLS_PER_SECOND = 410.0

_STATUS_EXC = {
    1: lambda: TypeError("'NoneType' object is not subscriptable"),          # opt_path is None (:174)
    2: lambda: ZeroDivisionError("float division by zero"),                   # (:270, :281)
    3: lambda: KeyError("time bin"),                                          # (:123-127)
    4: lambda: IndexError("uniform stream exhausted"),
    5: lambda: OverflowError("path longer than path_cap / chain_cap"),
}


class RRT:
    """RRT planner over the Catalina obstacle map, exploring without a goal (see `exploring`)."""

    def __init__(self, boundary, obstacles, sharkGrid, cell_list, exp_rate=1, dist_to_end=2, diff_max=0.5, freq=30):
        self.boundary_poly = boundary
        self.obstacle_list = obstacles
        self.mps_list = []
        self.time_bin = {}
        self.last_path = []
        self.exp_rate = exp_rate
        self.dist_to_end = dist_to_end
        self.diff_max = diff_max
        self.freq = freq
        self.cell_list = cell_list
        self.sharkGrid = sharkGrid
        self.precision = "f32"       # fast build for planning; "f64" = verification build
        self.device = 0
        self.sleep_between_replans = False
        self._envs = _world.EnvCache()
        self.t_start = 0.0

    # ------------------------------------------------------------------ marshalling
    def _env(self, obstacles, habitats):
        circles = _world.circles_of(obstacles)
        ring = _world.ring_of(self.boundary_poly)
        habs = _world.circles_of(habitats)
        return self._envs.get(circles, ring, habs, _world.grid_fingerprint(self.sharkGrid),
                              lambda: _world.grid_of(self.sharkGrid), device=self.device)

    @staticmethod
    def _row_to_mps(r):
        return Motion_plan_state(float(r[0]), float(r[1]), theta=float(r[2]), v=float(r[3]),
                                 traj_time_stamp=float(r[4]), length=float(r[5]))

    # ------------------------------------------------------------------ planners
    def exploring(self, initial, habitats, plot_interval, bin_interval, v, shark_interval, traj_time_stamp=False,
                  max_plan_time=5, max_traj_time=200.0, plan_time=True, weights=[-1, -1, -1], *,
                  iterations=None, seed=None, replicas=1, steer="arc", dubins_rho=1.0, dubins_eta=20.0,
                  near_radius=15.0, dubins_w=12):
        """reference :92-176.  Returns {"path length": float, "path": [list[MPS], {(t0,t1): list[MPS]}],
        "cost": [sum, [c0, c1, c2]]}.

        steer="dubins" (an extra of this build, NOT reference behaviour: the reference's Dubins steer is a
        commented-out call, :238-251) grows the tree with six-word Dubins edges toward sampled states and picks, among
        the nearest node and the nodes within near_radius, the parent with the cheapest path cost (planner mode 3)."""
        # plan_time & traj_time_stamp: time-bin pick (:122-127); plan_time only: pick by wall-clock
        # plan_time_stamp (get_closest_mps_time, :129-132) replayed on a simulated clock that spends
        # max_plan_time evenly over the steer calls; neither: nearest node to a random state (:136-139)
        mode = (0 if traj_time_stamp else 2) if plan_time else 1
        if steer == "dubins":
            mode = 3
        elif steer != "arc":
            raise ValueError("steer must be 'arc' or 'dubins'")
        iters = int(iterations) if iterations is not None else max(1, int(math.ceil(max_plan_time * REFERENCE_STEER_CALLS_PER_SECOND)))
        seed = random.getrandbits(63) if seed is None else int(seed)
        env = self._env(self.obstacle_list, habitats)
        pp = api.plan_params(iters, mode=mode, bin_interval=bin_interval, v=v, max_traj_time=max_traj_time,
                             dist_to_end=self.dist_to_end, diff_max=self.diff_max, freq=self.freq, min_dist=0.5,
                             weights=weights, chain_cap=255, path_cap=0, max_plan_time=max_plan_time,
                             dubins_rho=dubins_rho, dubins_eta=dubins_eta, near_radius=near_radius, dubins_w=dubins_w)
        start = [initial.x, initial.y, initial.theta, initial.traj_time_stamp, initial.length]
        R = max(1, int(replicas))
        starts = np.tile(np.array(start, dtype=np.float64), (R, 1))
        seeds = (np.arange(R, dtype=np.uint64) + np.uint64(seed)) & np.uint64(0x7FFFFFFFFFFFFFFF)
        r = api.plan_batch(env, starts, seeds, pp, self.precision)
        rec = r["records"]
        ok = rec["status"] == 0
        if not ok.any():
            raise _STATUS_EXC.get(int(rec["status"][0]), lambda: RuntimeError("planner failed"))()
        best = int(np.argmin(np.where(ok, rec["cost"][:, 0], np.inf)))        # strict <: first minimum wins
        # re-create the optimal path from its chain of stream positions (the planner stores no waypoints)
        # rows per edge: the parent node object + at most ceil(freq) arc waypoints (dubins_w - 1 Dubins samples)
        pp.path_cap = (max(int(math.ceil(self.freq)), int(dubins_w)) + 2) * (int(rec["depth"][best]) + 1) + 2
        path_rows, n_path = api.materialize(env, starts[best:best + 1], seeds[best:best + 1], r["chain"][best:best + 1],
                                            rec["depth"][best:best + 1], pp, self.precision)
        rows = path_rows[0, :n_path[0]]
        path = [self._row_to_mps(x) for x in rows]
        path[0] = initial                                   # the root of the course is the caller's object
        self.mps_list = path
        split = self.splitPath(path, shark_interval, [initial.traj_time_stamp, max_traj_time])
        c = rec["cost"][best]
        return {"path length": float(rec["path_length"][best]), "path": [path, split],
                "cost": [float(c[0]), [float(c[1]), float(c[2]), float(c[3])]]}

    def replanning(self, start, habitats, plan_time_budget, traj_time_length, replan_time_interval, weight):
        """reference :51-90: receding-horizon loop over `exploring`.  `initial` is detached from any
        previous tree (the reference sometimes walks on into the old tree): a segment whose
        start point was a NODE of the previous tree is costed without that tree's older waypoints.
        Deterministic replay: set `self.replan_seed` (inner call k uses seed + k) and `self.replan_iterations`."""
        replan_seed = getattr(self, "replan_seed", None)
        replan_iterations = getattr(self, "replan_iterations", None)
        traj = [start]
        time_dict = {}
        final_traj_time = list(self.sharkGrid.keys())[-1][1]
        plan_time = plan_time_budget + replan_time_interval
        count = 1
        oriHabitats = habitats.copy()
        while (traj[-1].traj_time_stamp + plan_time) < final_traj_time:
            if traj_time_length + traj[-1].traj_time_stamp > final_traj_time:
                traj_time_length = final_traj_time - traj[-1].traj_time_stamp
            temp = self.exploring(traj[-1], habitats, 0.5, 5, 2, plan_time, traj_time_stamp=True,
                                  max_plan_time=plan_time_budget,
                                  max_traj_time=(traj_time_length + traj[-1].traj_time_stamp), plan_time=True,
                                  weights=weight, iterations=replan_iterations,
                                  seed=None if replan_seed is None else replan_seed + count - 1)
            temp_path = temp["path"][1][list(temp["path"][1].keys())[0]]
            traj.extend(temp_path)
            time_dict[count] = [temp_path, habitats.copy()]
            habitats = self.removeHabitat(habitats, temp_path)
            count += 1
            if self.sleep_between_replans:
                import time
                time.sleep(replan_time_interval)
        cost = habitat_shark_cost_func(traj[1:], traj[-1].traj_time_stamp, oriHabitats, self.sharkGrid, weight=[-3, -3, -4])
        return [traj[1:], time_dict, cost]

    def planning(self, bin_interval=5, v=1, traj_time_stamp=False, max_plan_time=5, max_traj_time=200.0, plan_time=True):
        """reference :179-235 reads self.start / self.goal, which RRT.__init__ never sets: the method is
        dead there (AttributeError on the first line) and is kept dead here."""
        raise AttributeError("'RRT' object has no attribute 'start'")

    # ------------------------------------------------------------------ single-edge twins of the kernels
    def steer(self, mps, dist_to_end, diff_max, freq, min_dist, velocity=1, traj_time_stamp=False):
        """reference :237-295.  Draws from `random` in the reference's order, evaluates on the GPU."""
        u = [random.random()]
        n_expand = math.floor((0 + (freq - 0) * u[0]) / 1)
        for _ in range(n_expand):
            a, b = random.random(), random.random()
            u += [a, b]
            dist = 0 + (dist_to_end - 0) * a
            diff = -diff_max + (diff_max - (-diff_max)) * b
            if abs(dist) > abs(diff):
                u.append(random.random())
        parent = [[mps.x, mps.y, mps.theta, mps.traj_time_stamp, mps.length]]
        leaf, counts, wp, used, status = api.steer_arc(parent, u, [0, len(u)], [dist_to_end, diff_max, freq, min_dist, velocity],
                                                       "f64", self.device, wp_cap=max(int(n_expand), 1))
        if status[0] != 0:
            raise _STATUS_EXC[int(status[0])]()
        new = Motion_plan_state(float(leaf[0, 0]), float(leaf[0, 1]), theta=float(leaf[0, 2]),
                                traj_time_stamp=float(leaf[0, 3]), length=float(leaf[0, 4]))
        new.path = [mps] + [self._row_to_mps(r) for r in wp[0, :counts[0] - 1]]
        return new

    def check_collision(self, mps, obstacleList):
        """reference :530-549; True = safe"""
        if mps is None:
            return False
        env = self._env(obstacleList, [])
        pts = np.array([[p.x, p.y] for p in mps.path], dtype=np.float64).reshape(-1, 2)
        r = api.collide(env, [pts], "f64")[0]
        if r == 255:
            raise ValueError("min() arg is an empty sequence")
        return bool(r)

    def check_collision_obstacle(self, mps, obstacleList):
        """reference :551-556"""
        env = self._env(obstacleList, [])
        return bool(api.collide_points(env, [[mps.x, mps.y]], "f64")[0])

    def get_closest_mps(self, ran_mps, mps_list):
        """reference :505-513"""
        tree = np.array([[m.x, m.y] for m in mps_list], dtype=np.float64)
        return mps_list[int(api.nn(tree, [[ran_mps.x, ran_mps.y]], "f64", self.device)[0])]

    def get_random_mps(self, size_max=15):
        """reference :333-343"""
        x_min, y_min, x_max, y_max = self.boundary_poly.bounds
        return Motion_plan_state(random.uniform(x_min, x_max), random.uniform(y_min, y_max),
                                 theta=random.uniform(-math.pi, math.pi), size=random.uniform(0, size_max))

    def get_distance_angle(self, start_mps, end_mps):
        dx, dy = end_mps.x - start_mps.x, end_mps.y - start_mps.y
        return math.sqrt(dx ** 2 + dy ** 2), math.atan2(dy, dx)

    def generate_final_course(self, mps):
        """reference :321-331"""
        path = [mps]
        while mps.parent is not None:
            path.extend(reversed(mps.path))
            mps = mps.parent
        return path

    def angle_wrap(self, ang):
        while ang > math.pi:
            ang -= 2 * math.pi
        while ang < -math.pi:
            ang += 2 * math.pi
        return ang

    def cal_length(self, path):
        return sum(math.sqrt((path[i].x - path[i - 1].x) ** 2 + (path[i].y - path[i - 1].y) ** 2)
                   for i in range(1, len(path)))

    def splitPath(self, path, shark_interval, traj_time):
        """reference :590-602"""
        n_expand = math.floor(traj_time[1] / shark_interval)
        start = traj_time[0]
        res = {(start + i * shark_interval, start + (i + 1) * shark_interval): [] for i in range(n_expand)}
        for point in path:
            for t, arr in res.items():
                if t[0] <= point.traj_time_stamp <= t[1]:
                    arr.append(point)
                    break
        return res

    def removeHabitat(self, habitats, path):
        """reference :604-610 (mutates and returns its argument)"""
        for point in path:
            for habitat in habitats:
                if math.sqrt((point.x - habitat.x) ** 2 + (point.y - habitat.y) ** 2) <= habitat.size:
                    habitats.remove(habitat)
                    break
        return habitats


def createSharkGrid(filepath, cell_list):
    """reference :612-630: CSV rows 'time bin,grid' -> {(t0, t1): {cell_list[i].bounds: p_i}}"""
    out = {}
    with open(filepath, newline="") as f:
        for row in csv.DictReader(f):
            t0, t1 = (int(s) for s in row["time bin"].strip("()").split(", "))
            vals = row["grid"][1:-1].split(", ")
            out[(t0, t1)] = {cell_list[i].bounds: float(p) for i, p in enumerate(vals)}
    return out
