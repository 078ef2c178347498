"""TEST INFRASTRUCTURE ONLY -- golden vectors for the shark occupancy / AUV-detection grid builder
 from the UNMODIFIED reference:
SharkOccupancyGrid.convert (/root/reference/path_planning/sharkOccupancyGrid.py:47-71) under oracle/shims.

Run in the build container:  python oracle/make_golden_occupancy.py  -> tests/golden/occupancy.npz
Cells are the 10 m lattice clipped to the Catalina boundary (column-major; the reference's own
splitCell needs shapely.ops.split, parity unpinned -- the cell polygons are explicit inputs)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import harness as H  # noqa: E402


def main():
    ref = H.load_reference()
    saved = list(sys.path)
    sys.path[:0] = [H.SHIMS, os.path.join(H.REFERENCE_ROOT, "path_planning"), H.REFERENCE_ROOT]
    import sharkOccupancyGrid as sog
    sys.path[:] = saved
    world = json.load(open(os.path.join(ROOT, "tests", "golden", "catalina_map.json")))
    bpts = [tuple(p) for p in world["boundary"]]
    polys = H.lattice_cell_polygons(bpts, 10.0)
    boundary = ref.Polygon(bpts)
    cells = [ref.Polygon(p) for p in polys]
    MPS = ref.MPS
    rs = np.random.RandomState(77)
    out = {"cell_off": np.cumsum([0] + [len(p) for p in polys]).astype(np.int64),
           "cell_xy": np.concatenate([np.array(p) for p in polys]), "bounds": np.array(boundary.bounds)}
    cases = []
    # case 0: 5 sharks, 230 s, noisy straight tracks (some leave the boundary, some sit on lattice lines)
    # case 1: 2 sharks, short tracks, different bin width / range
    for ci, (S, n, dt, bi, rng_) in enumerate([(5, 230, 1.0, 50, 50), (2, 90, 0.5, 20, 25)]):
        tracks = []
        shark = {}
        for sid in range(1, S + 1):
            x0, y0 = rs.uniform(-330, -80), rs.uniform(-60, 90)
            vx, vy = rs.uniform(-0.5, 0.5, 2)
            pts = []
            for k in range(1, n + 1):
                t = k * dt
                x = x0 + vx * t + rs.normal(0, 1.5)
                y = y0 + vy * t + rs.normal(0, 1.5)
                if k % 37 == 0:                       # exactly on a lattice line / vertex of the cell grid
                    x = boundary.bounds[0] + 10.0 * round((x - boundary.bounds[0]) / 10.0)
                if k % 53 == 0:
                    y = boundary.bounds[1] + 10.0 * round((y - boundary.bounds[1]) / 10.0)
                pts.append((x, y, t))
            if sid == S:                              # last shark ends earlier: createBinList uses the longest
                pts = pts[: n // 2]
            tracks.append(np.array(pts))
            shark[sid] = [MPS(p[0], p[1], traj_time_stamp=p[2]) for p in pts]
        g = sog.SharkOccupancyGrid(10, boundary, bi, rng_, cells)
        arr, cellres = g.convert(shark)
        keys = list(arr.keys())
        grid = np.array([arr[k] for k in keys], dtype=np.float64)
        out["c%d_toff" % ci] = np.cumsum([0] + [len(t) for t in tracks]).astype(np.int64)
        out["c%d_trk" % ci] = np.concatenate(tracks)
        out["c%d_params" % ci] = np.array([10.0, bi, rng_], dtype=np.float64)
        out["c%d_grid" % ci] = grid
        out["c%d_bins" % ci] = np.array(keys, dtype=np.float64)
        # resultCell: {cell.bounds: value} without zeros, in cell_list order
        out["c%d_cellcount" % ci] = np.array([len(cellres[k]) for k in keys])
        out["c%d_cellvals" % ci] = np.concatenate([np.array(list(cellres[k].values())) for k in keys])
        cases.append((grid.shape, float(grid.sum())))
        print("case", ci, grid.shape, grid.sum(), [len(cellres[k]) for k in keys])
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "occupancy.npz"), **out)


if __name__ == "__main__":
    main()
