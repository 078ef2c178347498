"""TEST INFRASTRUCTURE ONLY (oracle shim for shapely.geometry) -- see shapely/__init__.py."""
from fractions import Fraction

_EPS = 2.0 ** -53
_ERRBOUND = (3.0 + 16.0 * _EPS) * _EPS  # Shewchuk's ccwerrboundA


def orient2d_sign(ax, ay, bx, by, cx, cy):
    """Exact sign of det[[ax-cx, ay-cy],[bx-cx, by-cy]] (float filter, Fraction fallback)."""
    detl = (ax - cx) * (by - cy)
    detr = (ay - cy) * (bx - cx)
    det = detl - detr
    if detl > 0.0:
        if detr <= 0.0:
            return (det > 0) - (det < 0)
        detsum = detl + detr
    elif detl < 0.0:
        if detr >= 0.0:
            return (det > 0) - (det < 0)
        detsum = -detl - detr
    else:
        detsum = abs(detr)
    if abs(det) > _ERRBOUND * detsum:
        return (det > 0) - (det < 0)
    F = Fraction
    d = (F(ax) - F(cx)) * (F(by) - F(cy)) - (F(ay) - F(cy)) * (F(bx) - F(cx))
    return (d > 0) - (d < 0)


class _Exterior:
    def __init__(self, pts):
        self.coords = list(pts) + [pts[0]]

    @property
    def xy(self):
        return ([p[0] for p in self.coords], [p[1] for p in self.coords])


class Polygon:
    def __init__(self, pts=None, holes=None):
        pts = [(float(p[0]), float(p[1])) for p in (pts or [])]
        if len(pts) > 1 and pts[0] == pts[-1]:
            pts = pts[:-1]
        self._pts = pts
        self.exterior = _Exterior(pts) if pts else None

    @property
    def bounds(self):
        xs = [p[0] for p in self._pts]
        ys = [p[1] for p in self._pts]
        return (min(xs), min(ys), max(xs), max(ys))

    def _locate(self, px, py):
        """+1 strictly inside, 0 on the boundary, -1 outside (exact)."""
        pts = self._pts
        n = len(pts)
        inside = False
        for i in range(n):
            ax, ay = pts[i]
            bx, by = pts[(i + 1) % n]
            if px == ax and py == ay:
                return 0
            if ay == py and by == py:
                if min(ax, bx) <= px <= max(ax, bx):
                    return 0
                continue
            if (ay > py) != (by > py):
                s = orient2d_sign(ax, ay, bx, by, px, py)
                if s == 0:
                    return 0
                # edge crosses the horizontal ray to +x iff p is left of the upward edge
                if (s > 0) == (by > ay):
                    inside = not inside
        return 1 if inside else -1

    @property
    def centroid(self):
        """GEOS Centroid for an area: signed-area-weighted mean of the fan triangles (base = first
        vertex) of the closed ring.  Ulp-level agreement with a particular GEOS build is not claimed
        (parity unpinned); callers treat the centroid as an input (astar_fixLenSOG.py:191)."""
        ring = self.exterior.coords
        bx, by = ring[0]
        area2_sum = 0.0
        cx3 = cy3 = 0.0
        for i in range(len(ring) - 1):
            (x1, y1), (x2, y2) = ring[i], ring[i + 1]
            a2 = (x1 - bx) * (y2 - by) - (x2 - bx) * (y1 - by)
            cx3 += a2 * (bx + x1 + x2)
            cy3 += a2 * (by + y1 + y2)
            area2_sum += a2
        return Point(cx3 / 3.0 / area2_sum, cy3 / 3.0 / area2_sum)

    def contains(self, other):
        return other.within(self)

    def touches(self, other):
        """boundary contact only (used as cell.touches(point) by sharkOccupancyGrid.py:226)"""
        if hasattr(other, "x") and hasattr(other, "y"):
            return self._locate(float(other.x), float(other.y)) == 0
        raise NotImplementedError


class Point:
    def __init__(self, x, y=None):
        if y is None:
            x, y = x
        self.x = float(x)
        self.y = float(y)

    @property
    def coords(self):
        return [(self.x, self.y)]

    def within(self, poly):
        return poly._locate(self.x, self.y) > 0


def box(minx, miny, maxx, maxy):
    return Polygon([(minx, miny), (maxx, miny), (maxx, maxy), (minx, maxy)])


class _Inert:
    def __init__(self, *a, **k):
        self.args = a


class MultiPolygon(_Inert):
    pass


class GeometryCollection(_Inert):
    pass


class LinearRing(_Inert):
    pass


class LineString(_Inert):
    pass
