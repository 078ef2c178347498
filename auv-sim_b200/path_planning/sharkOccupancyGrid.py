"""Drop-in for path_planning/sharkOccupancyGrid.py of auv-sim: shark tracks -> per-time-bin AUV-detection
probability grids (`SharkOccupancyGrid.convert`, /root/reference/path_planning/sharkOccupancyGrid.py:47-71),
the producer of the `{(t0, t1): {cell.bounds: p}}` dictionaries the planner's cost function reads.

The histogram / disc-sum arithmetic runs on the GPU (auvrrt_occupancy_grid, fp64, bit-identical to the
reference); there is no CPU fallback.  `splitCell` here clips the 10 m lattice to the boundary without
shapely (the reference uses shapely.ops.split, whose piece order is not reproducible offline: cells
are emitted column by column, bottom to top -- the order is part of the data).
"""
import math

import numpy as np

import _world
from _world import api


class _Ring:
    def __init__(self, pts):
        self.coords = list(pts) + [pts[0]]

    @property
    def xy(self):
        return [p[0] for p in self.coords], [p[1] for p in self.coords]


class Cell:
    """one clipped lattice cell: `.bounds` (what the cost function keys on) and `.exterior`"""

    def __init__(self, pts):
        self.pts = [(float(x), float(y)) for x, y in pts]
        xs, ys = [p[0] for p in self.pts], [p[1] for p in self.pts]
        self.bounds = (min(xs), min(ys), max(xs), max(ys))
        self.exterior = _Ring(self.pts)


def _clip(poly, x0, y0, x1, y1):
    """Sutherland-Hodgman clip of a polygon against an axis-aligned rectangle"""
    def run(pts, inside, cut):
        out = []
        for i, cur in enumerate(pts):
            prv = pts[i - 1]
            if inside(cur):
                if not inside(prv):
                    out.append(cut(prv, cur))
                out.append(cur)
            elif inside(prv):
                out.append(cut(prv, cur))
        return out

    def cut_x(xc):
        return lambda p, q: (xc, p[1] + (q[1] - p[1]) * (xc - p[0]) / (q[0] - p[0]))

    def cut_y(yc):
        return lambda p, q: (p[0] + (q[0] - p[0]) * (yc - p[1]) / (q[1] - p[1]), yc)

    pts = list(poly)
    for inside, cut in ((lambda p: p[0] >= x0, cut_x(x0)), (lambda p: p[0] <= x1, cut_x(x1)),
                        (lambda p: p[1] >= y0, cut_y(y0)), (lambda p: p[1] <= y1, cut_y(y1))):
        if not pts:
            break
        pts = run(pts, inside, cut)
    return pts


def splitCell(polygon, cell_size):
    """cells of the `cell_size` lattice anchored at the polygon's (minx, miny), clipped to the polygon
    (reference :376-393, without shapely)"""
    ring = [tuple(p) for p in _world.ring_of(polygon)]
    xs, ys = [p[0] for p in ring], [p[1] for p in ring]
    minx, miny, maxx, maxy = min(xs), min(ys), max(xs), max(ys)
    cells = []
    for i in range(int(math.ceil((maxx - minx) / cell_size))):
        for j in range(int(math.ceil((maxy - miny) / cell_size))):
            x0, y0 = minx + i * cell_size, miny + j * cell_size
            piece = _clip(ring, x0, y0, x0 + cell_size, y0 + cell_size)
            dedup = []
            for p in piece:
                if not dedup or p != dedup[-1]:
                    dedup.append((float(p[0]), float(p[1])))
            if len(dedup) > 1 and dedup[0] == dedup[-1]:
                dedup.pop()
            if len(dedup) >= 3:
                area = 0.5 * sum(dedup[k - 1][0] * dedup[k][1] - dedup[k][0] * dedup[k - 1][1] for k in range(len(dedup)))
                if abs(area) > 1e-9:
                    cells.append(Cell(dedup))
    return cells


def _cell_vertices(cell):
    if hasattr(cell, "pts"):
        return cell.pts
    ext = cell.exterior
    pts = [tuple(c[:2]) for c in ext.coords] if hasattr(ext, "coords") else list(zip(*ext.xy))
    if len(pts) > 1 and pts[0] == pts[-1]:
        pts = pts[:-1]
    return [(float(x), float(y)) for x, y in pts]


class SharkOccupancyGrid:
    def __init__(self, cell_size, boundary, bin_interval, detect_range, cell_list=None):
        self.cell_size = cell_size
        self.cell_list = cell_list if cell_list else splitCell(boundary, cell_size)
        self.bin_interval = bin_interval
        self.detect_range = detect_range
        self.boundary = boundary

    def _bounds(self):
        if hasattr(self.boundary, "bounds"):
            return tuple(float(v) for v in self.boundary.bounds)
        ring = _world.ring_of(self.boundary)
        return (ring[:, 0].min(), ring[:, 1].min(), ring[:, 0].max(), ring[:, 1].max())

    def createBinList(self):
        """reference :307-320"""
        longest = 0
        for traj in self.data.values():
            if traj[-1].traj_time_stamp > longest:
                longest = traj[-1].traj_time_stamp
        return [(i * self.bin_interval, (i + 1) * self.bin_interval)
                for i in range(math.floor(longest / self.bin_interval))]

    def cellToIndex(self, cell):
        minx, miny = self._bounds()[:2]
        lowx, lowy = cell.bounds[:2]
        return int((lowy - miny) / self.cell_size), int((lowx - minx) / self.cell_size)

    def convert(self, shark_dict, keep_zero_cells=False):
        """-> (resultArr {bin: 2-D list}, resultCell {bin: {cell.bounds: p}}), reference :47-71.

        Like the reference's convert2DArr (:284-292) the per-bin dicts drop zero-probability cells, so different bins
        generally list different cells.  The planner / cost entry points need every bin to list the same cells in the
        same order (their first-match scan depends on dict order, see _world.grid_of): keep_zero_cells=True (an extra
        of this build) keeps every cell of cell_list in every bin, in cell_list order -- the layout createSharkGrid
        (rrt_dubins.py:612-630) produces from the reference's CSV files."""
        self.data = shark_dict
        self.bin_list = self.createBinList()
        polys = [_cell_vertices(c) for c in self.cell_list]
        tracks = [np.array([[p.x, p.y, p.traj_time_stamp] for p in traj], dtype=np.float64).reshape(-1, 3)
                  for traj in shark_dict.values()]
        grid, _ = api.occupancy_grid(polys, self._bounds(), self.cell_size, self.bin_interval, self.detect_range, tracks)
        resultArr, resultCell = {}, {}
        idx = [self.cellToIndex(c) for c in self.cell_list]
        for b, key in enumerate(self.bin_list):
            resultArr[key] = grid[b].tolist()
            cells = {}
            for c, (row, col) in zip(self.cell_list, idx):                      # convert2DArr :284-292
                v = float(grid[b, row, col])
                if v != 0 or keep_zero_cells:
                    cells[c.bounds] = v
            resultCell[key] = cells
        return resultArr, resultCell


def getGridByTime(time, gridDict):
    for timebin, grid in gridDict.items():
        if timebin[0] <= time <= timebin[1]:
            return grid


def getGridByInterval(interval, gridDict):
    return {tb: g for tb, g in gridDict.items()
            if (tb[0] <= interval[0] <= tb[1]) or (tb[0] >= interval[0] and tb[1] <= interval[1])
            or (tb[0] <= interval[1] <= tb[1])}


# ---- raw shark tracks (data/sharkTrackingData.csv of auv-sim) ------------------------------------------
# No code in the reference reads that file (only a comment, robotSim.py:602); its layout is 32 sharks x 4
# rows (x, vx, y, vy) x 815 video frames at 30 fps (sharkTrajectory.py:33-37 stamps frame j with
# t = 0.03 j).  The tracks are in video pixels and cover 24 s, so using them with the Catalina planner
# needs an explicit map; CATALINA_TRACK_MAP is this build's definition (parity unpinned): the pixel box
# [370, 1334] x [4, 500] goes onto x in [-330, -80] m, y in [-60, 90] m and time is stretched x20 so the
# 815 frames span the planner's 500 s horizon.
CATALINA_TRACK_MAP = {"x0": 370.0, "y0": 4.0, "sx": 250.0 / 964.0, "sy": 150.0 / 496.0, "ox": -330.0, "oy": -60.0,
                      "frame_dt": 0.03, "time_scale": 20.0}


def load_shark_tracking_csv(path):
    """-> (x[S, N], y[S, N]) raw positions from a sharkTrackingData.csv-style file"""
    a = np.loadtxt(path, delimiter=",")
    return a[0::4], a[2::4]


def tracks_to_shark_dict(x, y, mapping=CATALINA_TRACK_MAP):
    """raw (x, y) tracks -> {shark_id: [Motion_plan_state(x, y, traj_time_stamp)]} in the map's frame"""
    from motion_plan_state import Motion_plan_state
    m = mapping
    out = {}
    for s in range(len(x)):
        out[s + 1] = [Motion_plan_state(m["ox"] + (float(x[s][j]) - m["x0"]) * m["sx"],
                                        m["oy"] + (float(y[s][j]) - m["y0"]) * m["sy"],
                                        traj_time_stamp=j * m["frame_dt"] * m["time_scale"]) for j in range(len(x[s]))]
    return out
