"""CPU: pin the C restatement of gym_rrt Planner_RRT (oracle/auvrrt_oracle.c, orc_gym_plan) against
tests/golden/gym_plan.npz, which the UNMODIFIED gym_rrt/envs/rrt_dubins.py produced on the sample
sequence (oracle/make_golden_gym.py).  Bit-exact."""
import math
import random

import numpy as np

from oracle import orc


def check_episode(ep, o):
    assert o["status"] == orc.OK
    assert o["steps"] == ep.steps and o["found"] == ep.found
    for k in ("parent", "nwp", "accepted", "done", "n_uniforms"):
        assert np.array_equal(o[k], getattr(ep, k)), (ep.name, k)
    assert np.array_equal(o["n_nodes_step"], ep.n_nodes) and np.array_equal(o["n_occupied_step"], ep.n_occupied)
    assert np.array_equal(o["cand"], ep.cand)
    assert np.array_equal(o["nodes"], ep.nodes)
    assert np.array_equal(o["occupied"], ep.flat(ep.occupied))
    assert o["n_path"] == len(ep.path) and np.array_equal(o["path"], ep.path)
    assert o["goal_arc_length"] == ep.goal_arc_length
    nz = np.flatnonzero(o["counts"])
    assert np.array_equal(np.stack([nz, o["counts"][nz]], 1), ep.counts_nz)


def test_py_hypot_matches_cpython():
    r = random.Random(5)
    for i in range(50000):
        a, b = r.uniform(-100, 100), r.uniform(-100, 100)
        if i % 3 == 0:
            a *= 10 ** r.uniform(-9, 9)
        assert orc.py_hypot(a, b) == math.hypot(a, b)
    assert orc.py_hypot(0.0, 0.0) == 0.0 and orc.py_hypot(3.0, -4.0) == 5.0


def test_gym_planning_bit_exact(gym_golden):
    n_found = 0
    for ep in gym_golden:
        w = orc.gym_world(ep.boundary, ep.obstacles, ep.goal, freq=ep.freq, cell_side_length=ep.cell_side,
                          subsections_in_cell=ep.subsections)
        assert orc.gym_grid_shape(w) == (ep.rows, ep.cols)
        acts = ep.flat_actions() if ep.actions is not None else None
        o = orc.gym_plan(w, ep.start, max_step=ep.max_step, seed=ep.seed, actions=acts)
        check_episode(ep, o)
        n_found += ep.found
    assert n_found >= 5


def test_gym_batch_matches_single(gym_golden):
    ep = gym_golden[0]
    w = orc.gym_world(ep.boundary, ep.obstacles, freq=10.0)
    rs = np.random.default_rng(3)
    Q = 64
    starts = np.column_stack([rs.uniform(5, 15, Q), rs.uniform(5, 15, Q), rs.uniform(-3, 3, Q)])
    goals = np.column_stack([rs.uniform(35, 45, Q), rs.uniform(35, 45, Q)])
    recs, status = orc.gym_plan_batch(w, starts, goals, np.arange(Q), max_step=100)
    assert (status == 0).all()
    for q in (0, 17, 63):
        w1 = orc.gym_world(ep.boundary, ep.obstacles, goals[q], freq=10.0)
        o = orc.gym_plan(w1, starts[q], max_step=100, seed=q)
        assert [o["steps"], int(o["found"]), o["n_nodes"]] == list(recs[q, :3])
