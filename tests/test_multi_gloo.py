"""CPU, world_size 2, gloo: the multi-GPU host logic (sharding, record gather, global min-cost
plan) with a stand-in local planner -- the CUDA planner itself is covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_plan(env, starts, seeds, params, precision):
    from auvrrt.api import RECORD_DTYPE
    rec = np.zeros(len(seeds), RECORD_DTYPE)
    rec["n_nodes"] = seeds.astype(np.int64) % 1000 + 1
    rec["cost"][:, 0] = -((seeds.astype(np.int64) * 7919) % 1013) / 100.0
    rec["status"] = (seeds.astype(np.int64) % 17 == 0).astype(np.int32)          # some queries fail
    rec["path_length"] = starts[:, 0]
    return {"records": rec}


def _worker(rank, world, port, Q, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "auv-sim_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from auvrrt import multi
    starts = np.zeros((Q, 5)); starts[:, 0] = np.arange(Q)
    seeds = np.arange(Q, dtype=np.uint64) + 3
    allrec, best = multi.plan_sharded(None, starts, seeds, None, "f32", plan_fn=_fake_plan)
    want = _fake_plan(None, starts, seeds, None, "f32")["records"]
    assert np.array_equal(allrec, want), "gathered records are not in global query order"
    c = np.where(want["status"] == 0, want["cost"][:, 0], np.inf)
    assert best == int(np.argmin(c)), (best, int(np.argmin(c)))               # lowest index among ties
    lo, hi = multi.shard_range(Q, rank, world)
    out[rank] = (lo, hi, best)
    dist.destroy_process_group()


@pytest.mark.parametrize("Q", [1, 2, 7, 64, 1001])
def test_sharded_plan_gather_world2(Q):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), Q, out), nprocs=world, join=True)
    ranges = [out[r][:2] for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == Q and ranges[0][1] == ranges[1][0]   # contiguous cover
    assert len({out[r][2] for r in range(world)}) == 1                                  # every rank agrees


def test_shard_range_properties():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "auv-sim_b200"))
    from auvrrt import multi
    for Q in (0, 1, 5, 4096, 1 << 20):
        for W in (1, 2, 3, 4, 8):
            r = [multi.shard_range(Q, k, W) for k in range(W)]
            assert r[0][0] == 0 and r[-1][1] == Q
            assert all(r[k][1] == r[k + 1][0] for k in range(W - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def _fake_gym(boundary, obstacles, starts, goals, seeds, max_step, **kw):
    from auvrrt.gym import GYM_RECORD_DTYPE
    rec = np.zeros(len(seeds), GYM_RECORD_DTYPE)
    rec["steps"] = seeds.astype(np.int64) % max_step + 1
    rec["done"] = (seeds.astype(np.int64) % 3 == 0)
    rec["cand"][:, 0] = starts[:, 0] + goals[:, 1]
    rec["n_uniforms"] = seeds.astype(np.int64) * 11
    return rec


def _fake_astar(env, q, **kw):
    from auvrrt.astar import ASTAR_RECORD_DTYPE
    rec = np.zeros(len(q), ASTAR_RECORD_DTYPE)
    rec["cost"] = -np.round(q["start"][:, 0] * 7.0) % 13          # many ties
    rec["status"] = (q["path_len_limit"] < 50).astype(np.int32)
    rec["n_expanded"] = q["start"][:, 1].astype(np.int32)
    return rec


def _worker_n1n4(rank, world, port, Q, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "auv-sim_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from auvrrt import multi
    from auvrrt.astar import ASTAR_QUERY_DTYPE
    rs = np.random.default_rng(Q)
    starts, goals = rs.uniform(0, 10, (Q, 3)), rs.uniform(30, 40, (Q, 2))
    seeds = np.arange(Q, dtype=np.uint64) + 5
    got = multi.gym_plan_sharded((0, 0, 50, 50), [], starts, goals, seeds, max_step=40, run_fn=_fake_gym)
    assert np.array_equal(got, _fake_gym(None, None, starts, goals, seeds, 40))
    q = np.zeros(Q, ASTAR_QUERY_DTYPE)
    q["start"] = rs.uniform(-100, 100, (Q, 2)); q["path_len_limit"] = rs.choice([40.0, 100.0, 300.0], Q)
    allrec, best = multi.astar_sharded(None, q, run_fn=_fake_astar)
    want = _fake_astar(None, q)
    assert np.array_equal(allrec, want)
    c = np.where(want["status"] == 0, want["cost"], np.inf)
    assert best == (int(np.argmin(c)) if np.isfinite(c).any() else -1)
    out[rank] = best
    dist.destroy_process_group()


@pytest.mark.parametrize("Q", [1, 9, 500])
def test_sharded_gym_and_astar_world2(Q):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_n1n4, args=(world, _free_port(), Q, out), nprocs=world, join=True)
    assert len({out[r] for r in range(world)}) == 1
