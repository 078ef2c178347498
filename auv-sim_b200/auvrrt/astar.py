"""Batched fixed-length lattice A* (path_planning/astar_fixLenSOG.py) on the GPU: libauvrrt.so, csrc/astar.cu.
One warp per query, fp64, bit-identical to the reference.  No CPU fallback."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import AstarRecord, check, lib

ASTAR_QUERY_DTYPE = np.dtype([("start", "<f8", (2,)), ("path_len_limit", "<f8"), ("weights", "<f8", (4,)), ("velocity", "<f8")])
ASTAR_RECORD_DTYPE = np.dtype([("status", "<i4"), ("n_expanded", "<i4"), ("n_nodes", "<i4"), ("n_path", "<i4"),
                               ("n_smooth", "<i4"), ("reserved", "<i4"), ("cost", "<f8"), ("path_len", "<f8")])
assert ASTAR_QUERY_DTYPE.itemsize == 64 and ASTAR_RECORD_DTYPE.itemsize == C.sizeof(AstarRecord) == 40


def _f64(a, shape):
    a = np.asarray(a, dtype=np.float64)
    return np.ascontiguousarray(a.reshape(shape)) if a.size else np.zeros([0 if s == -1 else s for s in shape])


def round_cells(cells):
    """cell bounds as get_cell_prob compares them (astar_fixLenSOG.py:494-495): Python's round(v, 2)"""
    return np.array([[round(float(v), 2) for v in c] for c in cells], dtype=np.float64).reshape(-1, 4)


def polygon_centroid(boundary):
    """area centroid of the boundary ring (what shapely's Polygon(...).centroid returns, :191): the
    signed-area-weighted mean of the fan triangles based at the first corner."""
    pts = [(float(p[0]), float(p[1])) for p in boundary]
    ring = pts + [pts[0]]
    bx, by = ring[0]
    area2 = cx3 = cy3 = 0.0
    for (x1, y1), (x2, y2) in zip(ring[:-1], ring[1:]):
        a2 = (x1 - bx) * (y2 - by) - (x2 - bx) * (y1 - by)
        cx3 += a2 * (bx + x1 + x2)
        cy3 += a2 * (by + y1 + y2)
        area2 += a2
    return (cx3 / 3.0 / area2, cy3 / 3.0 / area2)


class AstarEnv:
    def __init__(self, circles, boundary, habitats, bins, cells, probs, *, centroid=None, cells_are_rounded=False, device=0):
        """circles [K][3] obstacle_list; boundary [E][2] corners; habitats [H][3]; bins [T][2], cells [C][4]
        (bounds in dict order), probs [T][C]: the shark grid {(t0, t1): {cell.bounds: p}} flattened."""
        self.circles, self.boundary = _f64(circles, (-1, 3)), _f64(boundary, (-1, 2))
        self.habitats, self.bins = _f64(habitats, (-1, 3)), _f64(bins, (-1, 2))
        cells = _f64(cells, (-1, 4))
        self.cells_rounded = cells if cells_are_rounded else round_cells(cells)
        self.probs = _f64(probs, (len(self.bins), len(self.cells_rounded)))
        self.centroid = tuple(float(v) for v in (centroid if centroid is not None else polygon_centroid(self.boundary)))
        self.device = device
        self._h = C.c_void_p()
        p = lambda a: a.ctypes.data_as(_lib._dp)
        cen = (C.c_double * 2)(*self.centroid)
        check(lib().auvrrt_astar_env_create(p(self.circles), len(self.circles), p(self.boundary), len(self.boundary), cen,
                                            p(self.habitats), len(self.habitats), p(self.bins), len(self.bins),
                                            p(self.cells_rounded), len(self.cells_rounded), p(self.probs), int(device),
                                            C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().auvrrt_astar_env_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h


def make_queries(starts, path_len_limit=300.0, weights=(0, 10, 10, 100), velocity=1.0):
    starts = _f64(starts, (-1, 2))
    q = np.zeros(len(starts), ASTAR_QUERY_DTYPE)
    q["start"] = starts
    q["path_len_limit"] = path_len_limit
    q["weights"] = np.asarray(weights, dtype=np.float64)
    q["velocity"] = velocity
    return q


def astar_batch(env: AstarEnv, queries, *, node_cap=4096, path_cap=256, want_paths=True, trace=False):
    """-> dict(records, paths [Q][path_cap][6] (x, y, pathLen, time_stamp, cost, f; start -> goal), keep [Q][path_cap],
    and with trace: expand_order [Q][node_cap], node_xy [Q][node_cap][2])"""
    queries = np.ascontiguousarray(queries, dtype=ASTAR_QUERY_DTYPE)
    Q = len(queries)
    recs = np.zeros(Q, ASTAR_RECORD_DTYPE)
    paths = np.zeros((Q, path_cap, 6)) if want_paths else None
    keep = np.zeros((Q, path_cap), np.uint8) if want_paths else None
    order = np.zeros((Q, node_cap), np.int32) if trace else None
    xy = np.zeros((Q, node_cap, 2)) if trace else None
    check(lib().auvrrt_astar_batch(env.handle, queries.ctypes.data_as(C.c_void_p), C.c_int64(Q), int(node_cap), int(path_cap),
                                   recs.ctypes.data_as(C.c_void_p),
                                   paths.ctypes.data_as(_lib._dp) if want_paths else None,
                                   keep.ctypes.data_as(_lib._u8p) if want_paths else None,
                                   order.ctypes.data_as(_lib._i32p) if trace else None,
                                   xy.ctypes.data_as(_lib._dp) if trace else None))
    return {"records": recs, "paths": paths, "keep": keep, "expand_order": order, "node_xy": xy}
