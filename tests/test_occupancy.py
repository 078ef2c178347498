"""Shark occupancy / AUV-detection grid builder.
CPU: the C oracle against golden grids from the unmodified reference (bit-exact).
GPU: auvrrt_occupancy_grid and the drop-in SharkOccupancyGrid against the same grids (bit-exact)."""
import os
import sys

import numpy as np
import pytest

from oracle import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PP = os.path.join(ROOT, "auv-sim_b200", "path_planning")


@pytest.fixture(scope="module")
def golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "occupancy.npz"))
    polys = [z["cell_xy"][z["cell_off"][i]:z["cell_off"][i + 1]] for i in range(len(z["cell_off"]) - 1)]
    cases = []
    for ci in range(int(z["n_cases"])):
        toff, trk = z["c%d_toff" % ci], z["c%d_trk" % ci]
        cases.append({"tracks": [trk[toff[i]:toff[i + 1]] for i in range(len(toff) - 1)],
                      "params": z["c%d_params" % ci], "grid": z["c%d_grid" % ci], "bins": z["c%d_bins" % ci],
                      "cellcount": z["c%d_cellcount" % ci], "cellvals": z["c%d_cellvals" % ci]})
    return polys, z["bounds"], cases


def test_oracle_occupancy_bit_exact(golden):
    polys, bounds, cases = golden
    for c in cases:
        cs, bi, rg = c["params"]
        got = orc.occupancy(polys, bounds, cs, bi, rg, c["tracks"])
        assert got.shape == c["grid"].shape and np.array_equal(got, c["grid"])


def test_lattice_cells_match_fixture(golden, catalina_map):
    sys.path.insert(0, PP)
    for n in ("sharkOccupancyGrid", "_world", "motion_plan_state"):
        sys.modules.pop(n, None)
    import sharkOccupancyGrid as sog
    polys, bounds, _ = golden
    cells = sog.splitCell([tuple(p) for p in catalina_map["boundary"]], 10)
    assert len(cells) == len(polys) == 986
    assert np.allclose(np.array([c.bounds for c in cells]), np.array(catalina_map["cells"]), rtol=0, atol=1e-12)
    sys.path.remove(PP)


@pytest.mark.gpu
def test_gpu_occupancy_bit_exact(golden):
    from auvrrt import api
    polys, bounds, cases = golden
    for c in cases:
        cs, bi, rg = c["params"]
        grid, bins = api.occupancy_grid(polys, bounds, cs, bi, rg, c["tracks"])
        assert np.array_equal(grid, c["grid"])                      # histogram, disc sums, averages: bit-exact
        assert np.array_equal(bins, c["bins"])
    # size-independent properties at a larger size: 32 sharks x 815 samples (the shape of data/sharkTrackingData.csv)
    rs = np.random.RandomState(5)
    tracks = [np.stack([rs.uniform(-330, -80) + np.cumsum(rs.normal(0, 0.6, 815)), rs.uniform(-60, 90) + np.cumsum(rs.normal(0, 0.6, 815)),
                        0.03 * 30 * np.arange(1, 816)], 1) for _ in range(32)]
    grid, bins = api.occupancy_grid(polys, bounds, 10.0, 50.0, 50.0, tracks)
    want = orc.occupancy(polys, bounds, 10.0, 50.0, 50.0, tracks)
    assert np.array_equal(grid, want) and grid.shape[0] == len(bins) == 14
    assert grid.min() >= 0 and grid.max() <= 1.0 + 1e-12            # a detection probability


@pytest.mark.gpu
def test_gpu_dropin_shark_occupancy_grid(golden, catalina_map):
    sys.path.insert(0, PP)
    for n in ("sharkOccupancyGrid", "_world", "motion_plan_state", "rrt_dubins", "cost"):
        sys.modules.pop(n, None)
    import sharkOccupancyGrid as sog
    from motion_plan_state import Motion_plan_state as M
    polys, bounds, cases = golden
    boundary = [tuple(p) for p in catalina_map["boundary"]]
    cells = sog.splitCell(boundary, 10)
    c = cases[0]
    shark = {i + 1: [M(p[0], p[1], traj_time_stamp=p[2]) for p in t] for i, t in enumerate(c["tracks"])}
    g = sog.SharkOccupancyGrid(10, boundary, 50, 50, cells)
    arr, cell = g.convert(shark)
    assert list(arr.keys()) == [(0, 50), (50, 100), (100, 150), (150, 200)]
    assert np.array_equal(np.array([arr[k] for k in arr]), c["grid"])
    assert [len(cell[k]) for k in cell] == list(c["cellcount"])
    assert np.array_equal(np.concatenate([np.array(list(cell[k].values())) for k in cell]), c["cellvals"])
    # the grid feeds the planner's cost function unchanged
    import cost
    path = [M(-200.0 + i, 5.0, traj_time_stamp=10.0 * i) for i in range(15)]
    out = cost.habitat_shark_cost_func(path, 140.0, [], cell, [-3, -3, -4])
    assert out[0] < 0 and out[1][2] == out[0]
    sys.path.remove(PP)


@pytest.mark.gpu
def test_config3_grid_from_shark_tracking_data(golden, golden_dir, catalina_map):
    """BASELINE config 3, second half: the shark-occupancy cost built from data/sharkTrackingData.csv
    (frozen as tests/golden/shark_tracks_raw.npz) through the GPU grid builder, then planned with."""
    sys.path.insert(0, PP)
    for n in ("sharkOccupancyGrid", "_world", "motion_plan_state", "rrt_dubins", "cost"):
        sys.modules.pop(n, None)
    import sharkOccupancyGrid as sog
    import rrt_dubins
    from motion_plan_state import Motion_plan_state as M
    z = np.load(os.path.join(golden_dir, "shark_tracks_raw.npz"))
    shark = sog.tracks_to_shark_dict(z["x"].astype(np.float64), z["y"].astype(np.float64))
    assert len(shark) == 32 and len(shark[1]) == 815 and abs(shark[1][-1].traj_time_stamp - 488.4) < 1e-9
    boundary = [tuple(p) for p in catalina_map["boundary"]]
    cells = sog.splitCell(boundary, 10)
    arr, cellgrid = sog.SharkOccupancyGrid(10, boundary, 50, 50, cells).convert(shark)
    assert list(cellgrid.keys()) == [(50 * i, 50 * (i + 1)) for i in range(9)]
    polys, bounds, _ = golden
    tracks = [np.array([[p.x, p.y, p.traj_time_stamp] for p in t]) for t in shark.values()]
    want = orc.occupancy(polys, bounds, 10.0, 50.0, 50.0, tracks)
    assert np.array_equal(np.array([arr[k] for k in arr]), want)            # same grids as the oracle, bit for bit
    obstacles = [M(c[0], c[1], size=c[2]) for c in catalina_map["circles"]]
    habitats = [M(h[0], h[1], size=h[2]) for h in catalina_map["habitats"]]
    rrt = rrt_dubins.RRT(boundary, obstacles, cellgrid, cells)
    res = rrt.exploring(M(-200, 0), habitats, 0.5, 5, 2, 50, traj_time_stamp=True, max_traj_time=450,
                        weights=[-3, -3, -4], iterations=1024, seed=4, replicas=16)
    assert res["cost"][1][2] < 0 and res["path"][0][-1].traj_time_stamp >= 420      # the shark term is active
    sys.path.remove(PP)


@pytest.mark.gpu
def test_sparse_tracks_need_keep_zero_cells(golden, catalina_map):
    """per-bin dicts that list different cells (convert() drops zero-probability cells, like the reference's
    convert2DArr) cannot share the device cell index: the cost entry refuses such a grid with directions;
    keep_zero_cells=True always gives the planner-ready layout; and grids with equal keys but different probabilities
    are never confused (the world-model cache is keyed by content)."""
    sys.path.insert(0, PP)
    for n in ("sharkOccupancyGrid", "_world", "motion_plan_state", "rrt_dubins", "cost"):
        sys.modules.pop(n, None)
    import sharkOccupancyGrid as sog
    import cost
    from motion_plan_state import Motion_plan_state as M
    boundary = [tuple(p) for p in catalina_map["boundary"]]
    cells = sog.splitCell(boundary, 10)
    shark = {1: [M(-300.0 + 0.5 * i, 20.0, traj_time_stamp=2.0 * i) for i in range(60)],
             2: [M(-280.0, 40.0 - 0.5 * i, traj_time_stamp=2.0 * i) for i in range(60)]}
    g = sog.SharkOccupancyGrid(10, boundary, 50, 50, cells)
    arr, sparse = g.convert(shark)
    # (constructAUVGrid gives every cell inside the boundary a non-zero probability, so convert() itself rarely drops
    # one; a caller's own grid -- or a bin with an exactly-zero cell -- can.  Drop two cells from the second bin.)
    second = list(sparse.keys())[1]
    for c in list(sparse[second].keys())[3:5]:
        del sparse[second][c]
    path = [M(-300.0 + i, 20.0, traj_time_stamp=5.0 * i) for i in range(20)]
    with pytest.raises(NotImplementedError, match="keep_zero_cells"):
        cost.habitat_shark_cost_func(path, 95.0, [], sparse, [-3, -3, -4])
    arr2, dense = g.convert(shark, keep_zero_cells=True)
    assert all(list(v.keys()) == [c.bounds for c in cells] for v in dense.values())
    a = cost.habitat_shark_cost_func(path, 95.0, [], dense, [-3, -3, -4])
    assert a[1][2] < 0
    # a FRESH grid with the same bins and cells but doubled probabilities: twice the shark term, not a cached answer
    for k in range(6):
        doubled = {tb: {c: (k + 2) * p for c, p in v.items()} for tb, v in dense.items()}
        b = cost.habitat_shark_cost_func(path, 95.0, [], doubled, [-3, -3, -4])
        assert abs(b[1][2] - (k + 2) * a[1][2]) <= 1e-12 * abs(a[1][2]) * (k + 2)
    sys.path.remove(PP)
