"""Host-side marshalling shared by the drop-in modules: Python objects -> flat arrays -> auvrrt.Env."""
import os
import sys

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from auvrrt import api  # noqa: E402


def circles_of(objs):
    return np.array([[float(o.x), float(o.y), float(o.size)] for o in objs], dtype=np.float64).reshape(-1, 3)


def ring_of(boundary):
    """vertices of a shapely-like Polygon (`.exterior.xy` / `.exterior.coords`), a list of (x, y), or a
    list of objects with .x/.y; the closing vertex is dropped"""
    if boundary is None:
        return np.zeros((0, 2))
    if hasattr(boundary, "exterior"):
        ext = boundary.exterior
        if hasattr(ext, "xy"):
            xs, ys = ext.xy
            pts = list(zip(list(xs), list(ys)))
        else:
            pts = [tuple(c[:2]) for c in ext.coords]
    else:
        pts = [(p.x, p.y) if hasattr(p, "x") else (p[0], p[1]) for p in boundary]
    pts = [(float(a), float(b)) for a, b in pts]
    if len(pts) > 1 and pts[0] == pts[-1]:
        pts = pts[:-1]
    return np.array(pts, dtype=np.float64).reshape(-1, 2)


def grid_of(shark_dict):
    """{(t0, t1): {cell_bounds: p}} -> bins[T,2], cells[C,4], probs[T,C].

    The cost function scans each bin's dict in insertion order and takes the first matching cell
    (cost.py:181-184), so that order is part of the data: all bins must list the same cells in the
    same order (createSharkGrid, rrt_dubins.py:612-630, guarantees it)."""
    if not shark_dict:
        return np.zeros((0, 2)), np.zeros((0, 4)), np.zeros((0, 0))
    bins = np.array([[float(k[0]), float(k[1])] for k in shark_dict.keys()], dtype=np.float64)
    first = next(iter(shark_dict.values()))
    keys = list(first.keys())
    cells = np.array([[float(v) for v in k] for k in keys], dtype=np.float64).reshape(-1, 4)
    probs = np.zeros((len(bins), len(keys)))
    for i, g in enumerate(shark_dict.values()):
        if len(g) != len(keys) or (g is not first and list(g.keys()) != keys):
            raise NotImplementedError(
                "shark grid bins must list the same cells in the same order (bin %d lists %d cells, bin 0 lists %d). "
                "The first-match scan of cost.py:181-184 depends on each bin's own dict order, and the device cell "
                "index is built once for all bins.  createSharkGrid (rrt_dubins.py:612-630) always satisfies this; "
                "SharkOccupancyGrid.convert drops zero-probability cells per bin (as the reference's convert2DArr does): "
                "pass convert(..., keep_zero_cells=True) to get planner-ready grids." % (i, len(g), len(keys)))
        probs[i] = np.fromiter((float(p) for p in g.values()), dtype=np.float64, count=len(keys))
    return bins, cells, probs


class EnvCache:
    """auvrrt.Env handles keyed by the content of their inputs (small LRU)."""

    def __init__(self, cap=8):
        self.cap, self.items = cap, []

    def get(self, circles, ring, habitats, grid_id, grid_arrays_fn, device=0):
        key = (circles.tobytes(), ring.tobytes(), habitats.tobytes(), grid_id, device)
        for k, env in self.items:
            if k == key:
                return env
        bins, cells, probs = grid_arrays_fn()
        env = api.Env(circles, ring, habitats, bins, cells, probs, device=device)
        self.items.append((key, env))
        if len(self.items) > self.cap:
            _, old = self.items.pop(0)
            old.close()
        return env


def grid_fingerprint(shark_dict):
    """Content key of a shark grid {(t0, t1): {cell_bounds: p}}: a digest of its bins, cell bounds (in dict order) and
    probabilities.  Never the object's id(): CPython reuses the id of a freed dict, and a dict can be mutated in place,
    so an identity key silently served costs computed with a previous grid's probabilities."""
    if not shark_dict:
        return None
    import hashlib
    h = hashlib.blake2b(digest_size=16)
    for k, g in shark_dict.items():
        h.update(np.array([float(k[0]), float(k[1]), float(len(g))], dtype=np.float64).tobytes())
        h.update(np.fromiter((float(v) for b in g.keys() for v in b), dtype=np.float64, count=4 * len(g)).tobytes())
        h.update(np.fromiter((float(p) for p in g.values()), dtype=np.float64, count=len(g)).tobytes())
    return ("grid", h.hexdigest())
