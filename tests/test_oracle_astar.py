"""CPU: pin the C restatement of path_planning/astar_fixLenSOG.py (oracle/auvrrt_oracle.c, orc_astar)
against tests/golden/astar.npz, produced by the UNMODIFIED module (oracle/make_golden_astar.py).
The planner is deterministic and uses IEEE add / mul / sqrt only: bit-exact."""
import numpy as np

from oracle import orc


def oracle_world(world, n_bins=0):
    bins = world["bins"][:n_bins] if n_bins else world["bins"]
    probs = world["probs"][:n_bins] if n_bins else world["probs"]
    return orc.astar_world(world["circles"], world["boundary"], world["centroid"], world["habitats"], bins,
                           world["cells_rounded"], probs)


def check_case(case, o):
    assert o["status"] == case.want_status, (case.name, o["status"])
    exp_xy = o["node_xy"][o["expand_order"]]
    n = len(case.expanded)
    assert np.array_equal(exp_xy[:n], case.expanded[:, :2]), case.name
    if case.outcome == "ok":
        assert o["n_expanded"] == n + 1                 # + the node that met the length condition
        assert o["n_nodes"] == case.n_visited + 1       # + the start node, which is never marked visited
        assert np.array_equal(o["path"], case.nodes)
        assert o["cost"] == case.cost and np.array_equal(o["path"][::-1, 4], case.cost_list)
        assert np.array_equal(o["keep"], case.smooth_keep) and o["n_smooth"] == len(case.smooth_path)
        assert np.array_equal(o["path"][o["keep"] == 1][:, [0, 1, 3]], case.smooth_path)


def test_astar_bit_exact(astar_golden):
    world, cases = astar_golden
    assert sum(c.outcome == "ok" for c in cases) >= 8
    for case in cases:
        ow = oracle_world(world, case.n_bins)
        o = orc.astar(ow, case.start, case.velocity, case.limit, case.weights)
        check_case(case, o)


def test_astar_batch_matches_single(astar_golden):
    world, cases = astar_golden
    ow = oracle_world(world)
    ok = [c for c in cases if c.n_bins == 0]
    q = np.array([[c.start[0], c.start[1], c.limit, *c.weights, c.velocity] for c in ok])
    recs, cost, status = orc.astar_batch(ow, q)
    for i, c in enumerate(ok):
        assert status[i] == c.want_status
        if c.outcome == "ok":
            assert cost[i] == c.cost and recs[i, 2] == len(c.nodes) and recs[i, 3] == len(c.smooth_path)


def test_astar_overflow_and_helpers(astar_golden):
    world, cases = astar_golden
    ow = oracle_world(world)
    c = cases[1]
    assert orc.astar(ow, c.start, c.velocity, c.limit, c.weights, node_cap=50)["status"] == 5
    assert orc.astar(ow, c.start, c.velocity, c.limit, c.weights, path_cap=3)["status"] == 5
