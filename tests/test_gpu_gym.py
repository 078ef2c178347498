"""GPU: the vectorised gym_rrt Planner_RRT (csrc/gym.cu through the C ABI) against
 - tests/golden/gym_plan.npz, the unmodified gym_rrt/envs/rrt_dubins.py on the sample sequence:
   every step's decisions exactly (parent, kept primitives, accepted, done, uniforms consumed,
   occupied cells, node counts), floats to 1e-9 (fp64 build; device sin/cos/atan2 differ from glibc
   by an ulp or two);
 - the C oracle on a few thousand random episodes (fp64), and success statistics / path validity for
   the fp32 build (tolerance 1e-5 relative does not apply to a chaotic tree; validity does)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import orc  # noqa: E402
from oracle.harness import GYM_MAIN_OBSTACLES  # noqa: E402


@pytest.fixture(scope="module")
def gym():
    import auvrrt
    from auvrrt import gym as g
    assert auvrrt.api.device_count() > 0
    return g


def _batch(gym, ep, Q=3, precision=None, **kw):
    return gym.GymBatch(ep.boundary, ep.obstacles, Q, freq=ep.freq, cell_side_length=ep.cell_side,
                        subsections_in_cell=ep.subsections, node_cap=ep.max_step + 1, track_counts=True,
                        precision=gym.F64 if precision is None else precision, **kw)


def _reset(b, ep):
    b.reset(np.tile(ep.start, (b.Q, 1)), np.tile(ep.goal, (b.Q, 1)), np.full(b.Q, ep.seed, np.uint64))


def _check_final(b, ep, q):
    t = b.tree(q)
    assert np.allclose(t["nodes"], ep.nodes, rtol=1e-9, atol=1e-9)
    assert np.array_equal(t["occupied"], ep.flat(ep.occupied))
    cnt = b.counts(q, 1)[0]
    nz = np.flatnonzero(cnt)
    assert np.array_equal(np.stack([nz, cnt[nz]], 1), ep.counts_nz)
    if ep.found:
        p = b.path(q)
        assert p.shape == ep.path.shape
        assert np.allclose(p, ep.path, rtol=1e-9, atol=1e-9)


def test_golden_step_by_step(gym, gym_golden):
    for ep in gym_golden:
        b = _batch(gym, ep)
        assert (b.rows, b.cols) == (ep.rows, ep.cols)
        _reset(b, ep)
        acts = ep.flat(ep.actions) if ep.actions is not None else None
        n_calls = len(acts) if acts is not None else ep.steps
        s = 0
        for i in range(n_calls):
            if acts is None:
                r = b.plan(1, full_candidates=True)
            else:
                r = b.step(np.full(b.Q, acts[i], np.int32), full_candidates=True)
            assert (r["status"] == 0).all()
            assert all((r[k] == r[k][0]).all() for k in r.dtype.names if k != "cand")     # replicas agree
            r0 = r[1]
            if r0["last_parent"] < 0:               # empty cell: nothing happened
                assert acts is not None and r0["steps"] == s
                continue
            assert r0["steps"] == s + 1
            assert r0["last_parent"] == ep.parent[s] and r0["last_nwp"] == ep.nwp[s], (ep.name, s)
            assert r0["last_accepted"] == ep.accepted[s] and r0["done"] == ep.done[s], (ep.name, s)
            assert r0["n_nodes"] == ep.n_nodes[s] and r0["n_occupied"] == ep.n_occupied[s]
            assert r0["last_uniforms"] == ep.n_uniforms[s]
            assert np.allclose(r0["cand"], ep.cand[s], rtol=1e-9, atol=1e-9), (ep.name, s)
            rew = gym.GymBatch.rewards(r)[1]
            assert rew == (300 if ep.done[s] else (0 if ep.accepted[s] else -1))
            s += 1
        assert s == ep.steps
        if ep.found:
            assert r0["done"] == 1 and r0["n_path"] == len(ep.path)
            assert abs(r0["arc_length"] - ep.goal_arc_length) <= 1e-9 * max(1.0, ep.goal_arc_length)
            # a finished episode no longer steps
            r2 = b.plan(3)
            assert r2[0]["steps"] == ep.steps and r2[0]["done"] == 1 and r2[0]["n_path"] == len(ep.path)
        _check_final(b, ep, 2)
        b.close()


def test_golden_one_launch(gym, gym_golden):
    """Planner_RRT.planning in ONE launch gives the same tree as stepping (also with early exit on collisions)."""
    for ep in gym_golden:
        if ep.actions is not None:
            continue
        for full in (False, True):
            b = _batch(gym, ep)
            _reset(b, ep)
            r = b.plan(ep.max_step, full_candidates=full)[0]
            assert r["status"] == 0 and r["steps"] == ep.steps and bool(r["done"]) == ep.found
            assert r["n_nodes"] == len(ep.nodes) and r["n_uniforms"] == int(ep.n_uniforms.sum()) + ep.steps
            _check_final(b, ep, 0)
            b.close()


def _random_episodes(Q, seed):
    rs = np.random.default_rng(seed)
    starts = np.column_stack([rs.uniform(5, 15, Q), rs.uniform(5, 15, Q), rs.uniform(-np.pi, np.pi, Q)])
    goals = np.column_stack([rs.uniform(35, 45, Q), rs.uniform(35, 45, Q)])
    return starts, goals, np.arange(Q, dtype=np.uint64) + 1000


def test_random_episodes_vs_oracle_f64(gym):
    Q, max_step = 2048, 120
    starts, goals, seeds = _random_episodes(Q, 7)
    w = orc.gym_world((0, 0, 50, 50), GYM_MAIN_OBSTACLES, freq=10.0)
    want, st = orc.gym_plan_batch(w, starts, goals, seeds, max_step=max_step)
    assert (st == 0).all()
    b = gym.GymBatch((0, 0, 50, 50), GYM_MAIN_OBSTACLES, Q, freq=10.0, node_cap=max_step + 1, precision=gym.F64)
    b.reset(starts, goals, seeds)
    r = b.plan(max_step)
    assert (r["status"] == 0).all()
    same = (r["steps"] == want[:, 0]) & (r["done"] == want[:, 1]) & (r["n_nodes"] == want[:, 2]) & (r["n_occupied"] == want[:, 3])
    # an ulp of difference in sin/cos can flip a discrete decision in a rare episode
    assert same.sum() >= Q - 2, int((~same).sum())
    assert 0.2 < r["done"].mean() < 0.95
    b.close()


def _path_valid(path, start, goal, boundary, obstacles, tol):
    x, y = path[:, 0], path[:, 1]
    assert (x >= boundary[0] - tol).all() and (x <= boundary[2] + tol).all()
    assert (y >= boundary[1] - tol).all() and (y <= boundary[3] + tol).all()
    reff = np.maximum.accumulate(np.asarray(obstacles)[::-1, 2])[::-1]
    for (ox, oy, _), r in zip(obstacles, reff):
        assert (np.hypot(x - ox, y - oy) > r - tol).all()
    # ends at the start (the goal arc's first point when the tree is still only the root)
    assert np.hypot(*(path[-1, :2] - start[:2])) < 1e-3
    # the arc ends within exp_rate of the goal
    assert np.hypot(*(path[0, :2] - goal)) <= 1.0 + 1e-3
    # waypoints are at most dist_to_end apart along the tree part and exp_rate on the arc
    seg = np.hypot(np.diff(x), np.diff(y))
    assert seg.max() <= 2.0 + 1e-3


def test_fp32_build_statistics_and_paths(gym):
    Q, max_step = 8192, 120
    starts, goals, seeds = _random_episodes(Q, 11)
    w = orc.gym_world((0, 0, 50, 50), GYM_MAIN_OBSTACLES, freq=10.0)
    want, _ = orc.gym_plan_batch(w, starts, goals, seeds, max_step=max_step)
    b = gym.GymBatch((0, 0, 50, 50), GYM_MAIN_OBSTACLES, Q, freq=10.0, node_cap=max_step + 1, precision=gym.F32)
    b.reset(starts, goals, seeds)
    r = b.plan(max_step)
    assert (r["status"] == 0).all()
    # different uniforms (24-bit) -> different trees; the statistics must agree
    assert abs(r["done"].mean() - want[:, 1].mean()) < 0.03
    assert abs(r["steps"].mean() - want[:, 0].mean()) < 0.05 * want[:, 0].mean()
    assert abs(r["n_nodes"].mean() - want[:, 2].mean()) < 0.05 * want[:, 2].mean()
    done = np.flatnonzero(r["done"])[:64]
    for q in done:
        p = b.path(int(q))
        assert len(p) == r["n_path"][q]
        _path_valid(p, starts[q], goals[q], (0, 0, 50, 50), GYM_MAIN_OBSTACLES, 2e-3)
    b.close()


def test_vector_env_steps_with_agent_actions(gym):
    """RRTEnv.step semantics: the observation (dense counts) drives the next action."""
    Q = 256
    starts, goals, seeds = _random_episodes(Q, 3)
    b = gym.GymBatch((0, 0, 50, 50), GYM_MAIN_OBSTACLES, Q, freq=10.0, node_cap=64, track_counts=True, precision=gym.F32)
    b.reset(starts, goals, seeds)
    rs = np.random.default_rng(0)
    total_nodes = np.ones(Q, np.int64)
    for it in range(40):
        cnt = b.counts()
        assert np.array_equal(cnt.sum(1), total_nodes)
        # agent: a random occupied sub-cell per episode (argmax of noise over has_node)
        act = np.argmax((cnt > 0) * rs.random(cnt.shape), axis=1).astype(np.int32)
        if it % 7 == 3:
            act[::5] = b.n_subcells - 1                      # an empty cell: reward -1, nothing changes
        r = b.step(act)
        rew = gym.GymBatch.rewards(r)
        assert set(np.unique(rew)) <= {-1, 0, 300}
        live = r["status"] == 0
        assert live.all()
        total_nodes = r["n_nodes"].astype(np.int64)
    assert (r["steps"] > 20).all() or r["done"].any()
    b.close()


def test_errors(gym):
    from auvrrt._lib import AuvrrtError
    with pytest.raises(AuvrrtError):
        gym.GymBatch((0, 0, 50, 50), [], 4, cell_side_length=0.5)          # int(cell_side) == 0: ZeroDivisionError in the reference
    with pytest.raises(AuvrrtError):
        gym.GymBatch((0, 0, 50, 50), [], 0)
    b = gym.GymBatch((0, 0, 50, 50), [], 2, node_cap=3, precision=gym.F64)
    # start past the last row: never enters the grid, random.choice([]) raises -> KEY_ERROR
    b.reset([[10, 60, 0], [10, 10, 3.0]], [[40, 40], [10, 40]], [1, 2])
    r = b.plan(50)
    assert r["status"][0] == 3
    # episode 1 heads away from its goal: it keeps growing until node_cap overflows
    assert r["status"][1] in (0, 5)
    with pytest.raises(AuvrrtError):
        b.step(np.zeros(2, np.int32), records=True) if False else b._run(np.zeros(2, np.int32), 2, False, True)
    b.close()


def test_random_worlds_vs_oracle_f64(gym):
    """other boundaries (incl. one away from the origin), obstacle sets in non-monotone size order (the
    running-minimum quirk matters), cell sizes, subsection counts, exp_rate / dist_to_end / diff_max"""
    rs = np.random.default_rng(21)
    total = mism = 0
    for trial in range(10):
        x0, y0 = [(0.0, 0.0), (0.0, 0.0), (-40.0, -25.0), (10.0, 5.0)][trial % 4]
        W, Hh = rs.uniform(40, 120), rs.uniform(40, 120)
        boundary = (x0, y0, x0 + W, y0 + Hh)
        K = int(rs.integers(0, 12))
        obstacles = np.column_stack([rs.uniform(x0 + 0.2 * W, x0 + 0.8 * W, K), rs.uniform(y0 + 0.2 * Hh, y0 + 0.8 * Hh, K),
                                     rs.uniform(1.0, 7.0, K)]) if K else np.zeros((0, 3))
        kw = dict(exp_rate=float(rs.choice([1.0, 0.5, 2.0])), dist_to_end=float(rs.choice([2.0, 3.5])),
                  diff_max=float(rs.choice([0.5, 0.9])), freq=float(rs.choice([10.0, 25.0, 50.0])),
                  cell_side_length=float(rs.choice([2.0, 3.0, 5.0])), subsections_in_cell=int(rs.choice([4, 8, 12])))
        Q, max_step = 256, 80
        starts = np.column_stack([rs.uniform(x0 + 2, x0 + 0.15 * W, Q), rs.uniform(y0 + 2, y0 + 0.15 * Hh, Q), rs.uniform(-np.pi, np.pi, Q)])
        goals = np.column_stack([rs.uniform(x0 + 0.85 * W, x0 + W - 2, Q), rs.uniform(y0 + 0.85 * Hh, y0 + Hh - 2, Q)])
        seeds = rs.integers(0, 2**40, Q).astype(np.uint64)
        w = orc.gym_world(boundary, obstacles, **kw)
        want, st = orc.gym_plan_batch(w, starts, goals, seeds, max_step=max_step)
        b = gym.GymBatch(boundary, obstacles, Q, node_cap=max_step + 1, precision=gym.F64, **kw)
        b.reset(starts, goals, seeds)
        r = b.plan(max_step)
        assert np.array_equal(r["status"], st), trial
        ok = st == 0
        same = (r["steps"] == want[:, 0]) & (r["done"] == want[:, 1]) & (r["n_nodes"] == want[:, 2]) & (r["n_occupied"] == want[:, 3])
        total += int(ok.sum()); mism += int((~same & ok).sum())
        # one full tree per world, float for float
        q = int(np.flatnonzero(ok)[0])
        w1 = orc.gym_world(boundary, obstacles, goals[q], **kw)
        o = orc.gym_plan(w1, starts[q], max_step=max_step, seed=int(seeds[q]))
        if same[q]:
            t = b.tree(q)
            assert np.allclose(t["nodes"], o["nodes"], rtol=1e-9, atol=1e-9) and np.array_equal(t["parents"], o["parents"])
            assert np.array_equal(t["occupied"], o["occupied"]) and np.array_equal(t["cells"], o["node_cell"])
            if o["found"]:
                assert np.allclose(b.path(q), o["path"], rtol=1e-9, atol=1e-9)
        b.close()
    assert total > 2000 and mism <= 2, (total, mism)


def test_device_entry_points(gym):
    """auvrrt_gym_step_dev / auvrrt_gym_counts_dev on torch tensors give what the host-buffer calls give"""
    import torch
    from auvrrt._lib import check
    Q = 300
    starts, goals, seeds = _random_episodes(Q, 5)
    kw = dict(freq=10.0, node_cap=41, track_counts=True, precision=gym.F32)
    a = gym.GymBatch((0, 0, 50, 50), GYM_MAIN_OBSTACLES, Q, **kw)
    b = gym.GymBatch((0, 0, 50, 50), GYM_MAIN_OBSTACLES, Q, **kw)
    a.reset(starts, goals, seeds); b.reset(starts, goals, seeds)
    dev = torch.device("cuda", 0)
    d_recs = torch.zeros(Q * gym.GYM_RECORD_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    rs = np.random.default_rng(1)
    stream = torch.cuda.current_stream().cuda_stream
    for it in range(12):
        cnt = a.counts()
        act = np.argmax((cnt > 0) * rs.random(cnt.shape), axis=1).astype(np.int32)
        want = a.step(act)
        d_act = torch.from_numpy(act).to(dev)
        check(gym.lib().auvrrt_gym_step_dev(b.handle, d_act.data_ptr(), 1, 0, d_recs.data_ptr(), stream))
        torch.cuda.synchronize()
        got = np.frombuffer(d_recs.cpu().numpy().tobytes(), dtype=gym.GYM_RECORD_DTYPE)
        assert np.array_equal(got, want)
    # the resident counts array, read through the device pointer
    ptr = b.counts_device_ptr()
    assert ptr
    class _Dev:          # expose the library's device array to torch without a copy
        __cuda_array_interface__ = {"shape": (Q, b.n_subcells), "typestr": "<u2", "data": (int(ptr), False), "version": 3}
    torch.cuda.synchronize()
    host = torch.as_tensor(_Dev(), device=dev).cpu().numpy()
    assert np.array_equal(host, a.counts())
    # planning on the device entry point as well
    check(gym.lib().auvrrt_gym_step_dev(b.handle, None, 25, 0, d_recs.data_ptr(), stream))
    torch.cuda.synchronize()
    got = np.frombuffer(d_recs.cpu().numpy().tobytes(), dtype=gym.GYM_RECORD_DTYPE)
    assert np.array_equal(got, a.plan(25))
    a.close(); b.close()
