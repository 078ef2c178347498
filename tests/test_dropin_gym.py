"""Drop-in gym_rrt package (auv-sim_b200/gym_rrt): CPU part = signatures against the reference's,
plain classes, loud failure without a GPU; GPU part = Planner_RRT / RRTEnv / VecRRTEnv against the
golden episodes of the unmodified reference (tests/golden/gym_plan.npz)."""
import inspect
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "auv-sim_b200")
_NAMES = ("gym_rrt", "gym_rrt.envs", "gym_rrt.envs.rrt_env", "gym_rrt.envs.rrt_dubins",
          "gym_rrt.envs.motion_plan_state_rrt", "gym_rrt.envs.grid_cell_rrt")


@pytest.fixture(scope="module")
def g():
    saved = {n: sys.modules.pop(n) for n in _NAMES if n in sys.modules}
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    import gym_rrt.envs as envs
    import gym_rrt.envs.rrt_dubins as rd
    import gym_rrt.envs.motion_plan_state_rrt as mps
    import gym_rrt.envs.grid_cell_rrt as gc
    yield envs, rd, mps.Motion_plan_state, gc
    for n in _NAMES:
        sys.modules.pop(n, None)
    sys.modules.update(saved)


def test_signatures_match_reference(g):
    envs, rd, M, gc = g
    extras = ("seed", "replicas", "node_cap", "precision", "device", "track_counts", "max_steps")
    sig = lambda f: [p for p in inspect.signature(f).parameters if p not in extras]
    P = rd.Planner_RRT
    assert sig(P.__init__) == ["self", "start", "goal", "boundary", "obstacles", "habitats", "exp_rate", "dist_to_end",
                               "diff_max", "freq", "cell_side_length", "subsections_in_cell"]
    d = inspect.signature(P.__init__).parameters
    assert [d[k].default for k in ("exp_rate", "dist_to_end", "diff_max", "freq", "cell_side_length", "subsections_in_cell")] \
        == [1, 2, 0.5, 50, 2, 8]
    assert sig(P.planning) == ["self", "max_step", "min_length", "plan_time"]
    assert sig(P.generate_one_node) == ["self", "grid_cell", "step_num", "min_length"]
    assert sig(envs.RRTEnv.init_env) == ["self", "auv_init_pos", "shark_init_pos", "boundary_array", "grid_cell_side_length",
                                         "num_of_subsections", "obstacle_array", "habitat_grid"]
    assert sig(envs.RRTEnv.step) == ["self", "chosen_grid_cell_idx", "step_num"]
    m = M(1, 2)
    assert (m.z, m.theta, m.v, m.w, m.traj_time_stamp, m.plan_time_stamp, m.size, m.rl_state_id, m.parent, m.path, m.length) \
        == (0, 0, 0, 0, 0, 0, 0, None, None, [], 0)
    c = gc.Grid_cell_RRT(4, 6, side_length=2, num_of_subsections=8)
    th = [s.theta for s in c.subsection_cells]
    assert len(th) == 8 and th[0] == 0.0 and abs(th[3] - 3 * np.pi / 4) < 1e-12 and th[5] < 0      # wraps negative past pi
    assert c.delta_theta == float(2.0 * np.pi) / 8.0 and not c.has_node()


def test_no_gpu_fails_loudly(g):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    envs, rd, M, _ = g
    from auvrrt import AuvrrtError
    with pytest.raises(AuvrrtError, match="no CUDA device"):
        rd.Planner_RRT(M(10, 10), M(35, 40), [M(0, 0), M(50, 50)], [M(20, 20, size=3)], [])


def _objs(M, ep):
    start = M(float(ep.start[0]), float(ep.start[1]), theta=float(ep.start[2]))
    goal = M(float(ep.goal[0]), float(ep.goal[1]))
    boundary = [M(float(ep.boundary[0]), float(ep.boundary[1])), M(float(ep.boundary[2]), float(ep.boundary[3]))]
    obstacles = [M(float(o[0]), float(o[1]), size=float(o[2])) for o in ep.obstacles]
    return start, goal, boundary, obstacles


@pytest.mark.gpu
def test_planner_planning_matches_golden(g, gym_golden):
    envs, rd, M, _ = g
    n = 0
    for ep in gym_golden:
        if ep.actions is not None:
            continue
        start, goal, boundary, obstacles = _objs(M, ep)
        freq = int(ep.freq) if float(ep.freq).is_integer() else ep.freq
        pl = rd.Planner_RRT(start, goal, boundary, obstacles, [], freq=freq, cell_side_length=ep.cell_side,
                            subsections_in_cell=ep.subsections, seed=ep.seed, precision="f64")
        path, step, secs = pl.planning(max_step=ep.max_step)
        assert step == ep.steps and secs >= 0
        assert len(pl.mps_list) == len(ep.nodes)
        got = np.array([[m.x, m.y, m.theta, m.traj_time_stamp] for m in pl.mps_list])
        assert np.allclose(got, ep.nodes, rtol=1e-9, atol=1e-9)
        occ = np.array(pl.occupied_grid_cells_array, dtype=np.int64).reshape(-1, 3)
        assert np.array_equal(ep.flat(occ), ep.flat(ep.occupied))
        if ep.found:
            arr = np.array([[m.x, m.y, m.theta] for m in path])
            assert arr.shape == ep.path.shape and np.allclose(arr, ep.path, rtol=1e-9, atol=1e-9)
            assert abs(path[0].length - ep.goal_arc_length) < 1e-9 * max(1, ep.goal_arc_length)
            n += 1
        else:
            last_acc = bool(ep.accepted[-1])
            assert (path is not None) == last_acc
        # node_array views of the grid agree with the counts
        r, c, k = pl.occupied_grid_cells_array[0]
        assert pl.env_grid[r][c].subsection_cells[k].node_array[0] is pl.mps_list[0] or ep.name.startswith("plan_neg")
        assert pl.env_grid[r][c].has_node()
    assert n >= 5


@pytest.mark.gpu
def test_rrt_env_step_matches_golden(g, gym_golden):
    envs, rd, M, _ = g
    for ep in gym_golden:
        if ep.actions is None:
            continue
        start, goal, boundary, obstacles = _objs(M, ep)
        env = envs.RRTEnv()
        env.precision = "f64"
        env.seed(ep.seed)
        assert ep.freq == 10          # RRT_PLANNER_FREQ
        state = env.init_env(start, goal, boundary, ep.cell_side, ep.subsections, obstacles)
        nsub = ep.rows * ep.cols * ep.subsections
        assert state["rrt_grid"].shape == (nsub, 4) and state["has_node"].sum() == 1
        s = 0
        for a in ep.flat(ep.actions):
            before = state["rrt_grid_num_of_nodes_only"].copy()
            state, reward, done, info = env.step(int(a), step_num=s)
            if before[a] == 0:                                   # empty cell: -1, nothing changes
                assert reward == -1 and np.array_equal(before, state["rrt_grid_num_of_nodes_only"])
                continue
            assert reward == (300 if ep.done[s] else (0 if ep.accepted[s] else -1))
            assert done == bool(ep.done[s])
            assert state["rrt_grid_num_of_nodes_only"].sum() == ep.n_nodes[s]
            assert state["has_node"].sum() == ep.n_occupied[s]
            if ep.accepted[s] and not ep.done[s]:
                node = state["path"]
                assert np.allclose([node.x, node.y, node.theta, node.traj_time_stamp], ep.cand[s], rtol=1e-9, atol=1e-9)
            s += 1
        assert s == ep.steps
        nz = np.flatnonzero(state["rrt_grid_num_of_nodes_only"])
        assert np.array_equal(np.stack([nz, state["rrt_grid_num_of_nodes_only"][nz]], 1), ep.counts_nz)
        assert np.array_equal(state["rrt_grid"][:, 3], state["rrt_grid_num_of_nodes_only"])


@pytest.mark.gpu
def test_vec_env_and_replicas(g):
    envs, rd, M, _ = g
    from oracle.harness import GYM_MAIN_OBSTACLES
    Q = 512
    rs = np.random.default_rng(2)
    starts = np.column_stack([rs.uniform(5, 15, Q), rs.uniform(5, 15, Q), rs.uniform(-3, 3, Q)])
    goals = np.column_stack([rs.uniform(35, 45, Q), rs.uniform(35, 45, Q)])
    venv = envs.VecRRTEnv((0, 0, 50, 50), GYM_MAIN_OBSTACLES, Q, max_nodes=65)
    obs = venv.reset(starts, goals, np.arange(Q))
    assert obs.shape == (Q, venv.n_actions) and (obs.sum(1) == 1).all()
    ret = np.zeros(Q)
    for it in range(60):
        act = np.argmax((obs > 0) * rs.random(obs.shape), axis=1)
        obs, rew, done, recs = venv.step(act)
        ret += rew * (recs["status"] == 0)
        assert (obs.sum(1) == recs["n_nodes"]).all()
    assert done.any()
    q = int(np.flatnonzero(done)[0])
    p = venv.path(q)
    assert len(p) == recs["n_path"][q] and np.hypot(*(p[0, :2] - goals[q])) <= 1.0 + 1e-3
    venv.close()
    # replicas: many trees for one query, first finished one returned
    boundary = [M(0.0, 0.0), M(50.0, 50.0)]
    obstacles = [M(o[0], o[1], size=o[2]) for o in GYM_MAIN_OBSTACLES]
    pl = rd.Planner_RRT(M(10.0, 10.0, theta=0.0), M(35.0, 40.0), boundary, obstacles, [], seed=5, replicas=256)
    path, step, _ = pl.planning(max_step=200)
    assert isinstance(path, list) and step <= 200
    assert np.hypot(path[0].x - 35.0, path[0].y - 40.0) <= 1.0 + 1e-3
    assert abs(path[-1].x - 10.0) < 1e-3 and abs(path[-1].y - 10.0) < 1e-3


@pytest.mark.gpu
def test_rrt_env_long_episode_does_not_overflow(g):
    """the reference's RL drivers step an episode up to 1000 times (2000 at test time); the tree must hold that many
    nodes (round 1 capped it at 257 and raised OverflowError mid-episode), and a seeded env must not replay the same
    stream every episode"""
    envs, rd, M, _ = g
    env = envs.RRTEnv()
    env.seed(3)
    state = env.init_env(M(5.0, 5.0, theta=0.0), M(45.0, 45.0), [M(0.0, 0.0), M(50.0, 50.0)], 2, 8, [])
    rs = np.random.RandomState(0)
    steps = 0
    for s in range(700):
        occ = np.flatnonzero(state["has_node"])
        state, reward, done, _ = env.step(int(occ[rs.randint(len(occ))]), step_num=s)
        steps += 1
        if done:
            break
    assert steps > 300 or done
    first = state["rrt_grid_num_of_nodes_only"].copy()
    state = env.reset()
    occ = np.flatnonzero(state["has_node"])
    for s in range(40):
        occ = np.flatnonzero(state["has_node"])
        state, reward, done, _ = env.step(int(occ[0]), step_num=s)
    assert env._episode == 2
