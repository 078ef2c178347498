"""TEST INFRASTRUCTURE ONLY -- ctypes loader for the fp64 C oracle (oracle/auvrrt_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module, and only as the checker / CPU baseline.  The product (auv-sim_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liborc.so")

OK, NO_PATH, ZERO_DIV, KEY_ERROR, STREAM_END = 0, 1, 2, 3, 4

_dp = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "auvrrt_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH)):
        subprocess.run(["make", "-C", _HERE, "-B", "_build/liborc.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


class World(C.Structure):
    _fields_ = [("K", C.c_int), ("circles", _dp), ("E", C.c_int), ("poly", _dp),
                ("H", C.c_int), ("habitats", _dp), ("T", C.c_int), ("bins", _dp),
                ("C", C.c_int), ("cells", _dp), ("probs", _dp)]


class Stream(C.Structure):
    _fields_ = [("ext", _dp), ("n_ext", C.c_int64), ("key", C.c_uint64), ("f32u", C.c_int),
                ("pos", C.c_int64), ("exhausted", C.c_int)]


class SteerParams(C.Structure):
    _fields_ = [("dist_to_end", C.c_double), ("diff_max", C.c_double), ("freq", C.c_double),
                ("min_dist", C.c_double), ("velocity", C.c_double)]


class PlanParams(C.Structure):
    _fields_ = [("iterations", C.c_int), ("mode", C.c_int), ("bin_interval", C.c_double),
                ("v", C.c_double), ("max_traj_time", C.c_double), ("dist_to_end", C.c_double),
                ("diff_max", C.c_double), ("freq", C.c_double), ("min_dist", C.c_double),
                ("weights", C.c_double * 3), ("max_plan_time", C.c_double),
                ("dubins_rho", C.c_double), ("dubins_eta", C.c_double), ("near_radius", C.c_double),
                ("dubins_w", C.c_int)]


class Trace(C.Structure):
    _fields_ = [("parent", C.POINTER(C.c_int32)), ("safe", C.POINTER(C.c_uint8)),
                ("nwp", C.POINTER(C.c_int32)), ("leaf", _dp), ("upos", C.POINTER(C.c_int64)),
                ("cost_evals", _dp), ("n_cost_evals", C.c_int32), ("best_iter", C.c_int32),
                ("best_node", C.c_int32), ("n_nodes", C.c_int32), ("n_uniforms", C.c_int64),
                ("result", C.c_double * 5), ("path", _dp), ("path_cap", C.c_int32),
                ("n_path", C.c_int32), ("n_waypoints_total", C.c_int64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_stream_key.restype = C.c_uint64
        L.orc_stream_key.argtypes = [C.c_uint64]
        L.orc_stream_u.restype = C.c_double
        L.orc_stream_u.argtypes = [C.c_uint64, C.c_int64, C.c_int]
        L.orc_nn.restype = C.c_int
        L.orc_nn.argtypes = [_dp, _dp, C.c_int, C.c_double, C.c_double]
        L.orc_orient2d.restype = C.c_int
        L.orc_orient2d.argtypes = [C.c_double] * 6
        L.orc_point_in_polygon.restype = C.c_int
        L.orc_point_in_polygon.argtypes = [_dp, C.c_int, C.c_double, C.c_double]
        L.orc_check_collision.restype = C.c_int
        L.orc_check_collision.argtypes = [_dp, C.c_int, C.POINTER(World)]
        L.orc_check_collision_obstacle.restype = C.c_int
        L.orc_check_collision_obstacle.argtypes = [C.c_double, C.c_double, C.POINTER(World)]
        L.orc_cost_point.restype = C.c_double
        L.orc_dubins_shortest.restype = C.c_int
        L.orc_edge_dubins.restype = C.c_int
        L.orc_exploring.restype = C.c_int
        L.orc_steer_arc_ext.restype = C.c_int
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def _f64(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if shape is not None:
        a = a.reshape(shape)
    return a


class OracleWorld:
    """Flat fp64 arrays of the world model + the ctypes struct pointing at them."""

    def __init__(self, circles=(), boundary=(), habitats=(), bins=(), cells=(), probs=None):
        self.circles = _f64(circles, (-1, 3))
        self.boundary = _f64(boundary, (-1, 2))
        self.habitats = _f64(habitats, (-1, 3))
        self.bins = _f64(bins, (-1, 2))
        self.cells = _f64(cells, (-1, 4))
        T, Cn = len(self.bins), len(self.cells)
        self.probs = _f64(probs if probs is not None else np.zeros((T, Cn)), (T, Cn))
        self.c = World(len(self.circles), _p(self.circles), len(self.boundary), _p(self.boundary),
                       len(self.habitats), _p(self.habitats), T, _p(self.bins), Cn, _p(self.cells),
                       _p(self.probs))

    @classmethod
    def from_map(cls, world: dict, bins=(), probs=None, with_cells=True):
        return cls(world["circles"], world["boundary"], world["habitats"], bins,
                   world["cells"] if with_cells else (), probs)


def stream_u(seed, k, f32u=False):
    return lib().orc_stream_u(int(seed), int(k), int(f32u))


def nn(tree_xy, q):
    t = _f64(tree_xy, (-1, 2))
    tx, ty = np.ascontiguousarray(t[:, 0]), np.ascontiguousarray(t[:, 1])
    return lib().orc_nn(_p(tx), _p(ty), len(tx), float(q[0]), float(q[1]))


def steer_arc(parent, u, dist_to_end=2.0, diff_max=0.5, freq=30.0, min_dist=0.5, velocity=1.0):
    """-> (status, leaf[5], wp[n,6], n_uniforms_used)"""
    parent = _f64(parent, (5,))
    u = _f64(u, (-1,))
    sp = SteerParams(dist_to_end, diff_max, freq, min_dist, velocity)
    leaf = np.zeros(5)
    wp = np.zeros((int(np.ceil(freq)) + 2, 6))
    nwp = C.c_int(0)
    nused = C.c_int64(0)
    st = lib().orc_steer_arc_ext(_p(parent), _p(u), C.c_int64(len(u)), C.byref(sp), _p(leaf), _p(wp),
                                 C.byref(nwp), C.byref(nused))
    return st, leaf, wp[:nwp.value].copy(), nused.value


def point_in_polygon(poly, x, y):
    poly = _f64(poly, (-1, 2))
    return lib().orc_point_in_polygon(_p(poly), len(poly), float(x), float(y))


def check_collision(points_xy, world: OracleWorld):
    pts = _f64(points_xy, (-1, 2))
    return lib().orc_check_collision(_p(pts), len(pts), C.byref(world.c))


def check_collision_obstacle(x, y, world: OracleWorld):
    return lib().orc_check_collision_obstacle(float(x), float(y), C.byref(world.c))


def cost(points_xyt, total_traj_time, world: OracleWorld, weights, bin_mask=None):
    pts = _f64(points_xyt, (-1, 3))
    w = _f64(weights, (3,))
    out = np.zeros(4)
    mask = None
    if bin_mask is not None:
        mask = np.ascontiguousarray(np.asarray(bin_mask, dtype=np.uint8))
    lib().orc_cost(_p(pts), C.c_int(len(pts)), C.c_double(total_traj_time), C.byref(world.c),
                   mask.ctypes.data_as(C.POINTER(C.c_uint8)) if mask is not None else None,
                   _p(w), _p(out))
    return out


def cost_point(x, y, world: OracleWorld, visited, tb, weights):
    vis = np.ascontiguousarray(np.asarray(visited, dtype=np.uint8))
    w = _f64(weights, (3,))
    return lib().orc_cost_point(C.c_double(x), C.c_double(y), C.byref(world.c),
                                vis.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_int(tb), _p(w))


def plan_params(iterations, mode=0, bin_interval=5.0, v=2.0, max_traj_time=500.0, dist_to_end=2.0,
                diff_max=0.5, freq=30.0, min_dist=0.5, weights=(-3.0, -3.0, -4.0), max_plan_time=5.0,
                dubins_rho=1.0, dubins_eta=20.0, near_radius=15.0, dubins_w=12):
    return PlanParams(int(iterations), int(mode), bin_interval, v, max_traj_time, dist_to_end,
                      diff_max, freq, min_dist, (C.c_double * 3)(*[float(x) for x in weights]), float(max_plan_time),
                      float(dubins_rho), float(dubins_eta), float(near_radius), int(dubins_w))


def exploring(world: OracleWorld, start, params: PlanParams, seed=None, u=None, f32u=False,
              path_cap=4096):
    """-> dict(status, parent, safe, nwp, leaf, upos, cost_evals, result, path, ...)"""
    I = params.iterations
    start = _f64(start, (5,))
    parent = np.zeros(I, np.int32)
    safe = np.zeros(I, np.uint8)
    nwp = np.zeros(I, np.int32)
    leaf = np.zeros((I, 5))
    upos = np.zeros(I, np.int64)
    ce = np.zeros((I, 6))
    path = np.zeros((path_cap, 6))
    tr = Trace()
    tr.parent = parent.ctypes.data_as(C.POINTER(C.c_int32))
    tr.safe = safe.ctypes.data_as(C.POINTER(C.c_uint8))
    tr.nwp = nwp.ctypes.data_as(C.POINTER(C.c_int32))
    tr.leaf = _p(leaf)
    tr.upos = upos.ctypes.data_as(C.POINTER(C.c_int64))
    tr.cost_evals = _p(ce)
    tr.path = _p(path)
    tr.path_cap = path_cap
    if u is not None:
        u = _f64(u, (-1,))
        rng = Stream(_p(u), len(u), 0, 0, 0, 0)
    else:
        rng = Stream(None, 0, lib().orc_stream_key(int(seed)), int(f32u), 0, 0)
    st = lib().orc_exploring(C.byref(world.c), _p(start), C.byref(rng), C.byref(params), C.byref(tr))
    return {
        "status": st, "parent": parent, "safe": safe, "nwp": nwp, "leaf": leaf, "upos": upos,
        "cost_evals": ce[:tr.n_cost_evals].copy(), "best_iter": tr.best_iter,
        "best_node": tr.best_node, "n_nodes": tr.n_nodes, "n_uniforms": tr.n_uniforms,
        "result": np.array(list(tr.result)), "path": path[:min(tr.n_path, path_cap)].copy(),
        "n_path": tr.n_path, "n_waypoints_total": tr.n_waypoints_total,
    }


def exploring_batch(world: OracleWorld, starts, seeds, params: PlanParams, f32u=False, nthreads=0):
    starts = _f64(starts, (-1, 5))
    seeds = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64))
    Q = len(seeds)
    res = np.zeros((Q, 5))
    counts = np.zeros((Q, 3), np.int64)
    status = np.zeros(Q, np.int32)
    lib().orc_exploring_batch(C.byref(world.c), _p(starts), seeds.ctypes.data_as(C.POINTER(C.c_uint64)),
                              C.c_int(Q), C.byref(params), C.c_int(int(f32u)), _p(res),
                              counts.ctypes.data_as(C.POINTER(C.c_int64)),
                              status.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int(nthreads))
    return res, counts, status


def dubins_shortest(q0, q1, rho=1.0):
    """-> (word, params[3], length, all_len[6])"""
    q0, q1 = _f64(q0, (3,)), _f64(q1, (3,))
    prm = np.zeros(3)
    ln = C.c_double(0)
    al = np.zeros(6)
    wd = lib().orc_dubins_shortest(_p(q0), _p(q1), C.c_double(rho), _p(prm), C.byref(ln), _p(al))
    return wd, prm, ln.value, al


def dubins_sample(q0, rho, word, params, s):
    q0, params = _f64(q0, (3,)), _f64(params, (3,))
    q = np.zeros(3)
    lib().orc_dubins_sample(_p(q0), C.c_double(rho), C.c_int(word), _p(params), C.c_double(s), _p(q))
    return q


def edge_dubins(world: OracleWorld, q0, q1, rho, W):
    q0, q1 = _f64(q0, (3,)), _f64(q1, (3,))
    wp = np.zeros((W, 3))
    word = C.c_int(0)
    prm = np.zeros(3)
    ln = C.c_double(0)
    safe = lib().orc_edge_dubins(C.byref(world.c), _p(q0), _p(q1), C.c_double(rho), C.c_int(W), _p(wp),
                                 C.byref(word), _p(prm), C.byref(ln))
    return safe, word.value, prm, ln.value, wp


def edges_dubins_batch(world: OracleWorld, q0, q1, rho, W, nthreads=0):
    q0, q1 = _f64(q0, (-1, 3)), _f64(q1, (-1, 3))
    n = len(q0)
    safe = np.zeros(n, np.uint8)
    word = np.zeros(n, np.uint8)
    length = np.zeros(n)
    lib().orc_edges_dubins_batch(C.byref(world.c), _p(q0), _p(q1), C.c_int64(n), C.c_double(rho),
                                 C.c_int(W), safe.ctypes.data_as(C.POINTER(C.c_uint8)),
                                 word.ctypes.data_as(C.POINTER(C.c_uint8)), _p(length), C.c_int(nthreads))
    return safe, word, length


def edges_arc_batch(world: OracleWorld, parents, seeds, dist_to_end=2.0, diff_max=0.5, freq=30.0,
                    min_dist=0.5, velocity=2.0, f32u=False, nthreads=0):
    parents = _f64(parents, (-1, 5))
    seeds = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64))
    n = len(seeds)
    sp = SteerParams(dist_to_end, diff_max, freq, min_dist, velocity)
    safe = np.zeros(n, np.uint8)
    nwp = np.zeros(n, np.int32)
    leaf = np.zeros((n, 5))
    lib().orc_edges_arc_batch(C.byref(world.c), _p(parents), seeds.ctypes.data_as(C.POINTER(C.c_uint64)),
                              C.c_int64(n), C.byref(sp), C.c_int(int(f32u)),
                              safe.ctypes.data_as(C.POINTER(C.c_uint8)),
                              nwp.ctypes.data_as(C.POINTER(C.c_int32)), _p(leaf), C.c_int(nthreads))
    return safe, nwp, leaf


def num_threads():
    return lib().orc_num_threads()


def occupancy(cell_polys, bounds, cell_size, bin_interval, detect_range, tracks):
    """SharkOccupancyGrid.convert -> resultArr as an array [T, rows, cols].
    cell_polys: list of vertex lists; tracks: list (one per shark, dict order) of arrays [n,3] = x, y, t"""
    coff = np.zeros(len(cell_polys) + 1, np.int64)
    for i, p in enumerate(cell_polys):
        coff[i + 1] = coff[i] + len(p)
    cxy = _f64(np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1, 2) for p in cell_polys]), (-1, 2))
    toff = np.zeros(len(tracks) + 1, np.int64)
    for i, t in enumerate(tracks):
        toff[i + 1] = toff[i] + len(t)
    txy = _f64(np.concatenate([np.asarray(t, dtype=np.float64).reshape(-1, 3) for t in tracks]), (-1, 3))
    b = _f64(bounds, (4,))
    T, rows, cols = C.c_int(0), C.c_int(0), C.c_int(0)
    lib().orc_occupancy_dims(_p(b), C.c_double(cell_size), C.c_double(bin_interval), _p(txy),
                             toff.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int(len(tracks)), C.byref(T), C.byref(rows), C.byref(cols))
    out = np.zeros((max(T.value, 0), rows.value, cols.value))
    lib().orc_occupancy.restype = C.c_int
    r = lib().orc_occupancy(_p(cxy), coff.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int(len(cell_polys)), _p(b),
                            C.c_double(cell_size), C.c_double(bin_interval), C.c_double(detect_range), _p(txy),
                            toff.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int(len(tracks)), _p(out), C.c_int64(out.size))
    assert r == T.value
    return out


# ------------------------------------------------------------------ gym_rrt Planner_RRT ------
class GymWorld(C.Structure):
    _fields_ = [("x0", C.c_double), ("y0", C.c_double), ("x1", C.c_double), ("y1", C.c_double),
                ("K", C.c_int), ("circles", _dp), ("goal_x", C.c_double), ("goal_y", C.c_double),
                ("exp_rate", C.c_double), ("dist_to_end", C.c_double), ("diff_max", C.c_double),
                ("freq", C.c_double), ("cell_side", C.c_double), ("subsections", C.c_int)]


_ip = C.POINTER(C.c_int32)


def gym_world(boundary, obstacles, goal=(0.0, 0.0), exp_rate=1.0, dist_to_end=2.0, diff_max=0.5, freq=50.0,
              cell_side_length=2.0, subsections_in_cell=8):
    circles = _f64(obstacles, (-1, 3)) if len(obstacles) else np.zeros((0, 3))
    w = GymWorld(float(boundary[0]), float(boundary[1]), float(boundary[2]), float(boundary[3]),
                 len(circles), _p(circles), float(goal[0]), float(goal[1]), float(exp_rate), float(dist_to_end),
                 float(diff_max), float(freq), float(cell_side_length), int(subsections_in_cell))
    w._keep = circles
    return w


def gym_grid_shape(w: GymWorld):
    r, c = C.c_int(), C.c_int()
    lib().orc_gym_grid_shape(C.byref(w), C.byref(r), C.byref(c))
    return r.value, c.value


def py_hypot(a, b):
    f = lib().orc_py_hypot
    f.restype = C.c_double
    f.argtypes = [C.c_double, C.c_double]
    return f(a, b)


def gym_plan(w: GymWorld, start, max_step=200, seed=None, u=None, actions=None, path_cap=4096, want_counts=True):
    """Planner_RRT.planning (actions None) or the RRTEnv.step loop over flat sub-cell ids."""
    L = lib()
    L.orc_gym_plan.restype = C.c_int
    start = _f64(start, (3,))
    cap = max_step + 1
    rows, cols = gym_grid_shape(w)
    rec = np.zeros(6, np.int32)
    nodes = np.zeros((cap, 4))
    parents = np.zeros(cap, np.int32)
    node_cell = np.zeros(cap, np.int32)
    occupied = np.zeros(cap, np.int32)
    counts = np.zeros(rows * cols * w.subsections, np.int32) if want_counts else None
    trace = np.zeros((max_step, 11))
    path = np.zeros((path_cap, 3))
    arc_len = C.c_double()
    if u is not None:
        u = _f64(u, (-1,))
        rng = Stream(_p(u), len(u), 0, 0, 0, 0)
    else:
        rng = Stream(None, 0, L.orc_stream_key(int(seed)), 0, 0, 0)
    if actions is not None:
        actions = np.ascontiguousarray(np.asarray(actions, dtype=np.int32))
        assert len(actions) == max_step
    st = L.orc_gym_plan(C.byref(w), _p(start), C.byref(rng), C.c_int(max_step),
                        actions.ctypes.data_as(_ip) if actions is not None else None,
                        rec.ctypes.data_as(_ip), _p(nodes), parents.ctypes.data_as(_ip), node_cell.ctypes.data_as(_ip),
                        occupied.ctypes.data_as(_ip), counts.ctypes.data_as(_ip) if want_counts else None,
                        _p(trace), _p(path), C.c_int(path_cap), C.byref(arc_len))
    steps, found, n, n_occ, n_path, last = [int(v) for v in rec]
    tr = trace[:steps]
    return {
        "status": st, "steps": steps, "found": bool(found), "n_nodes": n, "last": last,
        "nodes": nodes[:n].copy(), "parents": parents[:n].copy(), "node_cell": node_cell[:n].copy(),
        "occupied": occupied[:n_occ].copy(), "counts": counts, "path": path[:min(n_path, path_cap)].copy(),
        "n_path": n_path, "goal_arc_length": arc_len.value, "n_uniforms_total": int(rng.pos),
        "parent": tr[:, 0].astype(np.int32), "nwp": tr[:, 1].astype(np.int32), "accepted": tr[:, 2].astype(np.uint8),
        "done": tr[:, 3].astype(np.uint8), "n_nodes_step": tr[:, 4].astype(np.int32),
        "n_occupied_step": tr[:, 5].astype(np.int32), "n_uniforms": tr[:, 6].astype(np.int32), "cand": tr[:, 7:11].copy(),
        "grid_shape": (rows, cols),
    }


def gym_plan_batch(w: GymWorld, starts, goals, seeds, max_step=200, nthreads=0):
    starts = _f64(starts, (-1, 3))
    goals = _f64(goals, (-1, 2))
    seeds = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64))
    Q = len(seeds)
    recs = np.zeros((Q, 6), np.int32)
    status = np.zeros(Q, np.int32)
    lib().orc_gym_plan_batch(C.byref(w), _p(starts), _p(goals), seeds.ctypes.data_as(C.POINTER(C.c_uint64)),
                             C.c_int(Q), C.c_int(max_step), C.c_int(nthreads), recs.ctypes.data_as(_ip),
                             status.ctypes.data_as(_ip))
    return recs, status


# ------------------------------------------------------------------ lattice A* (astar_fixLenSOG) ---
class AstarWorld(C.Structure):
    _fields_ = [("K", C.c_int), ("circles", _dp), ("E", C.c_int), ("boundary", _dp), ("centroid", C.c_double * 2),
                ("H", C.c_int), ("habitats", _dp), ("T", C.c_int), ("bins", _dp), ("C", C.c_int), ("cells_r", _dp),
                ("probs", _dp), ("topn", _dp)]


def astar_world(circles, boundary, centroid, habitats, bins, cells_rounded, probs):
    """cells_rounded: cell bounds after Python's round(v, 2) (oracle/harness.py round_cells)."""
    L = lib()
    keep = dict(circles=_f64(circles, (-1, 3)) if len(circles) else np.zeros((0, 3)),
                boundary=_f64(boundary, (-1, 2)),
                habitats=_f64(habitats, (-1, 3)) if len(habitats) else np.zeros((0, 3)),
                bins=_f64(bins, (-1, 2)) if len(bins) else np.zeros((0, 2)),
                cells=_f64(cells_rounded, (-1, 4)) if len(cells_rounded) else np.zeros((0, 4)))
    T, Cc = len(keep["bins"]), len(keep["cells"])
    keep["probs"] = _f64(probs, (T, Cc)) if T and Cc else np.zeros((T, Cc))
    keep["topn"] = np.zeros((T, Cc + 1))
    L.orc_astar_topn(_p(keep["probs"]), C.c_int(T), C.c_int(Cc), _p(keep["topn"]))
    w = AstarWorld(len(keep["circles"]), _p(keep["circles"]), len(keep["boundary"]), _p(keep["boundary"]),
                   (C.c_double * 2)(float(centroid[0]), float(centroid[1])), len(keep["habitats"]), _p(keep["habitats"]),
                   T, _p(keep["bins"]), Cc, _p(keep["cells"]), _p(keep["probs"]), _p(keep["topn"]))
    w._keep = keep
    return w


def astar(w: AstarWorld, start, velocity=1.0, path_len_limit=300.0, weights=(0, 10, 10, 100), node_cap=4096, path_cap=512):
    L = lib()
    L.orc_astar.restype = C.c_int
    rec = np.zeros(4, np.int32)
    cost = C.c_double()
    path = np.zeros((path_cap, 6))
    keep = np.zeros(path_cap, np.uint8)
    order = np.zeros(node_cap, np.int32)
    xy = np.zeros((node_cap, 2))
    st = L.orc_astar(C.byref(w), _p(_f64(start, (2,))), C.c_double(velocity), C.c_double(path_len_limit),
                     _p(_f64(weights, (4,))), C.c_int(node_cap), C.c_int(path_cap), rec.ctypes.data_as(_ip),
                     C.byref(cost), _p(path), keep.ctypes.data_as(C.POINTER(C.c_uint8)), order.ctypes.data_as(_ip), _p(xy))
    n_exp, n, n_path, n_smooth = [int(v) for v in rec]
    ok = st == 0
    return {"status": st, "n_expanded": n_exp, "n_nodes": n, "n_path": n_path, "n_smooth": n_smooth, "cost": cost.value,
            "path": path[:n_path].copy() if ok else np.zeros((0, 6)), "keep": keep[:n_path].copy() if ok else np.zeros(0, np.uint8),
            "expand_order": order[:n_exp].copy(), "node_xy": xy[:n].copy()}


def astar_batch(w: AstarWorld, queries, node_cap=4096, path_cap=512, nthreads=0):
    """queries [Q][8] = start x, y, path_len_limit, w1, w2, w3, w4, velocity"""
    queries = _f64(queries, (-1, 8))
    Q = len(queries)
    recs = np.zeros((Q, 4), np.int32)
    cost = np.zeros(Q)
    status = np.zeros(Q, np.int32)
    lib().orc_astar_batch(C.byref(w), _p(queries), C.c_int(Q), C.c_int(node_cap), C.c_int(path_cap), C.c_int(nthreads),
                          recs.ctypes.data_as(_ip), _p(cost), status.ctypes.data_as(_ip))
    return recs, cost, status
