"""GPU: the device-resident sharded planner (auvrrt.multi.plan_sharded_device) at world size 1 with the REAL planner:
NCCL all-gather of the records, the two MIN all-reduces for the global minimum-cost plan, the winner's path."""
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_plan_sharded_device_world1(catalina_map, shark_grid):
    import torch
    import torch.distributed as dist
    from auvrrt import api, multi
    assert api.device_count() > 0
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=0, world_size=1, device_id=dev)
    try:
        env = api.Env.from_map(catalina_map, shark_grid[0], shark_grid[1])
        Q = 300
        starts = np.tile([-200.0, 0.0, 0.0, 0.0, 0.0], (Q, 1))
        starts[:, 0] += np.linspace(-40, 40, Q)
        seeds = np.arange(Q, dtype=np.uint64) + 11
        for group in (32, 1):
            pp = api.plan_params(512, group=group, chain_cap=200)
            r = multi.plan_sharded_device(env, starts, seeds, pp, "f32", want_path=True)
            rec = r["records"].cpu().numpy().reshape(-1).view(api.RECORD_DTYPE)
            want = api.plan_batch(env, starts, seeds, pp, "f32")
            assert np.array_equal(rec, want["records"])                       # the gather returns the planner's records
            c = np.where(rec["status"] == 0, rec["cost"][:, 0], np.inf)
            assert r["best"] == int(np.argmin(c))                             # first minimum wins (strict <)
            # the broadcast path is the winner's materialised path
            b = r["best"]
            pp2 = api.plan_params(512, group=group, chain_cap=200, path_cap=34 * (int(rec["depth"][b]) + 1))
            rows, n_path = api.materialize(env, starts[b:b + 1], seeds[b:b + 1], want["chain"][b:b + 1], rec["depth"][b:b + 1], pp2, "f32")
            assert np.array_equal(r["path"], rows[0, :n_path[0]])
            assert abs(r["path"][-1, 5] - rec["path_length"][b]) <= 1e-4 * max(1.0, rec["path_length"][b])
            # the host-facing wrapper routes through the same path on NCCL
            allrec, best = multi.plan_sharded(env, starts, seeds, pp, "f32")
            assert best == r["best"] and np.array_equal(allrec, rec)
        env.close()
    finally:
        dist.destroy_process_group()
