"""GPU: the batched lattice A* (csrc/astar.cu through the C ABI) against tests/golden/astar.npz (the
unmodified path_planning/astar_fixLenSOG.py) and against the C oracle on a sweep of starts / limits /
weights.  Deterministic fp64 planner: everything bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import orc  # noqa: E402


@pytest.fixture(scope="module")
def astar():
    import auvrrt
    from auvrrt import astar as a
    assert auvrrt.api.device_count() > 0
    return a


def _env(astar, world, n_bins=0):
    bins = world["bins"][:n_bins] if n_bins else world["bins"]
    probs = world["probs"][:n_bins] if n_bins else world["probs"]
    return astar.AstarEnv(world["circles"], world["boundary"], world["habitats"], bins, world["cells_rounded"], probs,
                          centroid=world["centroid"], cells_are_rounded=True)


def test_golden_queries_bit_exact(astar, astar_golden):
    world, cases = astar_golden
    # the product's own centroid agrees with the fixture's
    assert np.allclose(astar.polygon_centroid(world["boundary"]), world["centroid"], rtol=0, atol=1e-12)
    full = [c for c in cases if c.n_bins == 0]
    env = _env(astar, world)
    q = np.zeros(len(full), astar.ASTAR_QUERY_DTYPE)
    for i, c in enumerate(full):
        q[i] = (tuple(c.start), c.limit, tuple(c.weights), c.velocity)
    r = astar.astar_batch(env, q, trace=True)
    for i, c in enumerate(full):
        rec = r["records"][i]
        assert rec["status"] == c.want_status, (c.name, rec["status"])
        n = len(c.expanded)
        exp_xy = r["node_xy"][i][r["expand_order"][i][:rec["n_expanded"]]]
        assert np.array_equal(exp_xy[:n], c.expanded[:, :2]), c.name
        if c.outcome == "ok":
            assert rec["n_expanded"] == n + 1 and rec["n_nodes"] == c.n_visited + 1
            p = r["paths"][i][:rec["n_path"]]
            assert np.array_equal(p, c.nodes), c.name
            assert rec["cost"] == c.cost and np.array_equal(p[::-1, 4], c.cost_list)
            k = r["keep"][i][:rec["n_path"]]
            assert np.array_equal(k, c.smooth_keep) and rec["n_smooth"] == len(c.smooth_path)
            assert np.array_equal(p[k == 1][:, [0, 1, 3]], c.smooth_path)
    for c in cases:
        if c.n_bins:
            env2 = _env(astar, world, c.n_bins)
            rec = astar.astar_batch(env2, astar.make_queries([c.start], c.limit, c.weights, c.velocity))["records"][0]
            assert rec["status"] == c.want_status
            env2.close()
    env.close()


def test_sweep_vs_oracle(astar, astar_golden):
    world, _ = astar_golden
    env = _env(astar, world)
    ow = orc.astar_world(world["circles"], world["boundary"], world["centroid"], world["habitats"], world["bins"],
                         world["cells_rounded"], world["probs"])
    rs = np.random.default_rng(4)
    Q = 192
    q = np.zeros(Q, astar.ASTAR_QUERY_DTYPE)
    q["start"] = np.round(np.column_stack([rs.uniform(-420, 40, Q), rs.uniform(-120, 150, Q)]), 2)
    q["path_len_limit"] = rs.choice([60.0, 100.0, 150.0, 220.0, 300.0], Q)
    q["weights"] = np.column_stack([np.zeros(Q), rs.choice([0.0, 1.0, 10.0], Q), rs.choice([0.0, 10.0], Q), rs.choice([0.0, 50.0, 100.0], Q)])
    q["velocity"] = rs.choice([1.0, 1.5, 2.0], Q)
    r = astar.astar_batch(env, q, trace=True)
    oq = np.column_stack([q["start"], q["path_len_limit"], q["weights"], q["velocity"]])
    recs, cost, status = orc.astar_batch(ow, oq)
    g = r["records"]
    assert np.array_equal(g["status"], status)
    ok = status == 0
    assert ok.sum() > Q // 3
    assert np.array_equal(g["n_expanded"], recs[:, 0]) and np.array_equal(g["n_nodes"], recs[:, 1])
    assert np.array_equal(g["n_path"][ok], recs[ok, 2]) and np.array_equal(g["n_smooth"][ok], recs[ok, 3])
    assert np.array_equal(g["cost"][ok], cost[ok])
    for i in np.flatnonzero(ok)[:12]:
        o = orc.astar(ow, q["start"][i], q["velocity"][i], q["path_len_limit"][i], q["weights"][i])
        assert np.array_equal(r["paths"][i][:g["n_path"][i]], o["path"])
        assert np.array_equal(r["keep"][i][:g["n_path"][i]], o["keep"])
        assert np.array_equal(r["expand_order"][i][:g["n_expanded"][i]], o["expand_order"])
    env.close()


def test_caps_and_errors(astar, astar_golden):
    from auvrrt._lib import AuvrrtError
    world, cases = astar_golden
    env = _env(astar, world)
    c = cases[1]
    q = astar.make_queries([c.start], c.limit, c.weights, c.velocity)
    assert astar.astar_batch(env, q, node_cap=50)["records"][0]["status"] == 5
    assert astar.astar_batch(env, q, path_cap=3)["records"][0]["status"] == 5
    rec = astar.astar_batch(env, q, want_paths=False)["records"][0]
    assert rec["status"] == 0 and rec["cost"] == c.cost and rec["n_path"] == len(c.nodes)
    with pytest.raises(AuvrrtError):
        astar.astar_batch(env, q, node_cap=8192)
    env.close()
    with pytest.raises(AuvrrtError):
        astar.AstarEnv(world["circles"], world["boundary"][:2], world["habitats"], world["bins"], world["cells_rounded"],
                       world["probs"], centroid=(0.0, 0.0), cells_are_rounded=True)


def test_random_worlds_vs_oracle(astar):
    """synthetic worlds: convex boundaries with 3..7 corners, random obstacles, lattice cells of other sizes with
    unrounded bounds, other time-bin widths -- still bit for bit"""
    rs = np.random.default_rng(33)
    n_ok = 0
    for trial in range(6):
        E = int(rs.integers(3, 8))
        ang = np.sort(rs.uniform(0, 2 * np.pi, E))
        R = rs.uniform(120, 200)
        boundary = np.column_stack([R * np.cos(ang), R * np.sin(ang)]) * rs.uniform(0.8, 1.0, (E, 1)) + rs.uniform(-50, 50, 2)
        K = int(rs.integers(0, 15))
        circles = np.column_stack([rs.uniform(-R, R, K), rs.uniform(-R, R, K), rs.uniform(2, 15, K)]) + np.array([boundary[:, 0].mean(), boundary[:, 1].mean(), 0]) if K else np.zeros((0, 3))
        H = int(rs.integers(0, 6))
        habitats = np.column_stack([rs.uniform(-R, R, H), rs.uniform(-R, R, H), rs.uniform(5, 30, H)]) if H else np.zeros((0, 3))
        cs = float(rs.choice([10.0, 7.5, 20.0]))
        x0, y0, x1, y1 = boundary[:, 0].min(), boundary[:, 1].min(), boundary[:, 0].max(), boundary[:, 1].max()
        xs, ys = np.arange(x0, x1, cs), np.arange(y0, y1, cs)
        cells = np.array([[x, y, min(x + cs, x1), min(y + cs, y1)] for x in xs for y in ys]) + rs.uniform(0, 1e-4, (len(xs) * len(ys), 4)) * 0
        cells_r = astar.round_cells(cells)
        T = int(rs.integers(3, 9)); bw = float(rs.choice([40.0, 50.0, 75.0]))
        bins = np.array([[t * bw, (t + 1) * bw] for t in range(T)])
        probs = rs.random((T, len(cells))) * (rs.random((T, len(cells))) < 0.3)
        cen = astar.polygon_centroid(boundary)
        env = astar.AstarEnv(circles, boundary, habitats, bins, cells_r, probs, centroid=cen, cells_are_rounded=True)
        ow = orc.astar_world(circles, boundary, cen, habitats, bins, cells_r, probs)
        Q = 48
        q = np.zeros(Q, astar.ASTAR_QUERY_DTYPE)
        c0 = np.array(cen)
        q["start"] = np.round(c0 + rs.uniform(-0.5 * R, 0.5 * R, (Q, 2)), 2)
        q["path_len_limit"] = rs.choice([50.0, 90.0, 130.0, 200.0], Q)
        q["weights"] = np.column_stack([np.zeros(Q), rs.choice([0.0, 3.0, 10.0], Q), rs.choice([0.0, 10.0], Q), rs.choice([1.0, 50.0, 100.0], Q)])
        q["velocity"] = rs.choice([0.8, 1.0, 2.0], Q)
        r = astar.astar_batch(env, q, trace=True)
        oq = np.column_stack([q["start"], q["path_len_limit"], q["weights"], q["velocity"]])
        recs, cost, status = orc.astar_batch(ow, oq)
        g = r["records"]
        assert np.array_equal(g["status"], status), trial
        assert np.array_equal(g["n_expanded"], recs[:, 0]) and np.array_equal(g["n_nodes"], recs[:, 1])
        ok = status == 0
        n_ok += int(ok.sum())
        assert np.array_equal(g["cost"][ok], cost[ok]) and np.array_equal(g["n_smooth"][ok], recs[ok, 3])
        for i in np.flatnonzero(ok)[:4]:
            o = orc.astar(ow, q["start"][i], q["velocity"][i], q["path_len_limit"][i], q["weights"][i])
            assert np.array_equal(r["paths"][i][:g["n_path"][i]], o["path"]) and np.array_equal(r["keep"][i][:g["n_path"][i]], o["keep"])
        env.close()
    assert n_ok > 60


def test_device_entry_point(astar, astar_golden):
    """auvrrt_astar_batch_dev on torch tensors == the host-buffer call"""
    import torch
    from auvrrt._lib import check
    world, cases = astar_golden
    env = _env(astar, world)
    ok = [c for c in cases if c.outcome == "ok"][:6]
    q = np.zeros(len(ok), astar.ASTAR_QUERY_DTYPE)
    for i, c in enumerate(ok):
        q[i] = (tuple(c.start), c.limit, tuple(c.weights), c.velocity)
    want = astar.astar_batch(env, q, node_cap=2048, path_cap=64)
    dev = torch.device("cuda", 0)
    Q = len(q)
    d_q = torch.from_numpy(q.view(np.uint8).reshape(Q, 64).copy()).to(dev)
    d_rec = torch.zeros(Q * 40, dtype=torch.uint8, device=dev)
    d_paths = torch.zeros((Q, 64, 6), dtype=torch.float64, device=dev)
    d_keep = torch.zeros((Q, 64), dtype=torch.uint8, device=dev)
    wsb = int(astar.lib().auvrrt_astar_workspace_bytes(Q, 2048))
    d_ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    check(astar.lib().auvrrt_astar_batch_dev(env.handle, d_q.data_ptr(), Q, 2048, 64, d_ws.data_ptr(), wsb, d_rec.data_ptr(),
                                             d_paths.data_ptr(), d_keep.data_ptr(), None, None,
                                             torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    rec = np.frombuffer(d_rec.cpu().numpy().tobytes(), dtype=astar.ASTAR_RECORD_DTYPE)
    assert np.array_equal(rec, want["records"])
    assert np.array_equal(d_paths.cpu().numpy(), want["paths"]) and np.array_equal(d_keep.cpu().numpy(), want["keep"])
    # too small a workspace is refused
    from auvrrt._lib import AuvrrtError
    with pytest.raises(AuvrrtError):
        check(astar.lib().auvrrt_astar_batch_dev(env.handle, d_q.data_ptr(), Q, 2048, 64, d_ws.data_ptr(), 1024, d_rec.data_ptr(),
                                                 d_paths.data_ptr(), d_keep.data_ptr(), None, None, None))
    env.close()
