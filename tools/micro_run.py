#!/usr/bin/env python
"""small driver for profiling the micro-kernels: python tools/micro_run.py [edges|nn] [n]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "auv-sim_b200"))
import bench  # noqa
from auvrrt import api, device as adev  # noqa

what = sys.argv[1] if len(sys.argv) > 1 else "edges"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
dev = torch.device("cuda", 0)
if what == "edges":
    rs = np.random.RandomState(1234)
    K = 500
    circles = np.stack([rs.uniform(-467.4, 82.4, K), rs.uniform(-153.5, 191.2, K), rs.uniform(1, 5, K)], 1)
    world, _, _ = bench.load_world()
    env4 = api.Env(circles=circles, boundary=world["boundary"])
    g = torch.Generator(device=dev); g.manual_seed(1)
    q0 = torch.stack([torch.rand(n, device=dev, generator=g) * 549.8 - 467.4, torch.rand(n, device=dev, generator=g) * 344.7 - 153.5,
                      (torch.rand(n, device=dev, generator=g) * 2 - 1) * np.pi], 1).contiguous()
    ang = (torch.rand(n, device=dev, generator=g) * 2 - 1) * np.pi
    dist_ = torch.rand(n, device=dev, generator=g) * 38 + 2
    q1 = torch.stack([q0[:, 0] + dist_ * torch.cos(ang), q0[:, 1] + dist_ * torch.sin(ang),
                      (torch.rand(n, device=dev, generator=g) * 2 - 1) * np.pi], 1).contiguous()
    safe = torch.zeros(n, dtype=torch.uint8, device=dev); word = torch.zeros(n, dtype=torch.uint8, device=dev)
    length = torch.zeros(n, device=dev)
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        adev.edges_dubins_dev(env4, q0, q1, 1.0, 20, safe, word, length, "f32")
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    import os
    flop = 6 * 20 * K + 6 * 20 * 5 + 140 + 24 * 20
    print("edges-dubins n", n, "edges/s %.4g" % (n / min(ts)), "algorithmic TFLOP/s %.2f" % (n * flop / min(ts) / 1e12), "safe", float(safe.float().mean()),
          {k: os.environ[k] for k in os.environ if k.startswith("AUVRRT_")})
elif what.startswith("mode"):
    # python tools/micro_run.py mode1|mode3 Q : the warp-per-tree planner in nearest-node mode / Dubins best-parent mode
    world, bins_, probs_ = bench.load_world()
    env = api.Env.from_map(world, bins_, probs_, device=0)
    pp = api.plan_params(bench.ITERS, mode=1) if what == "mode1" else api.plan_params(1024, mode=3, v=1.0, max_traj_time=200.0)
    st, sd = bench.make_queries(0, n)
    pl = adev.DevicePlanner(env, pp, "f32", n, want_chain=True)
    pl.set_queries(st, sd)
    for _ in range(3):
        pl.launch()
    torch.cuda.synchronize()
elif what == "tpt":
    # python tools/micro_run.py tpt Q : the thread-per-tree planner on Q queries (config-5 scale on one GPU)
    world, bins_, probs_ = bench.load_world()
    env = api.Env.from_map(world, bins_, probs_, device=0)
    pp = api.plan_params(bench.ITERS, group=1)
    st, sd = bench.make_queries(0, n)
    pl = adev.DevicePlanner(env, pp, "f32", n, want_chain=True)
    pl.set_queries(st, sd)
    ts = []
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pl.launch(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    print("tpt Q", n, "edges/s %.4g" % (n * bench.ITERS / min(ts)), "plans/s %.4g" % (n / min(ts)), ts)
elif what.startswith("catalina") or what.startswith("config4"):
    # python tools/micro_run.py catalina[-allpairs|-nocost] n : the thread-per-edge arc kernel at Catalina scale
    # python tools/micro_run.py config4[-allpairs|-nocost] n  : the same on config 4's 500 random circles (bench.py micro_config4_arc*)
    import os
    world, bins_, probs_ = bench.load_world()
    if what.startswith("config4"):
        rs = np.random.RandomState(1234)
        circles = np.stack([rs.uniform(-467.4, 82.4, 500), rs.uniform(-153.5, 191.2, 500), rs.uniform(1, 5, 500)], 1)
        env = api.Env(circles=circles, boundary=world["boundary"], habitats=world["habitats"], bins=bins_, cells=world["cells"],
                      probs=probs_, device=0)
        g = torch.Generator(device=dev); g.manual_seed(2)
        par = torch.stack([torch.rand(n, device=dev, generator=g) * 549.8 - 467.4, torch.rand(n, device=dev, generator=g) * 344.7 - 153.5,
                           (torch.rand(n, device=dev, generator=g) * 2 - 1) * np.pi, torch.rand(n, device=dev, generator=g) * 400.0,
                           torch.zeros(n, device=dev)], 1).contiguous()
    else:
        env = api.Env.from_map(world, bins_, probs_, device=0)
        par = bench.catalina_parents(world, n, dev)
    sd = torch.arange(n, device=dev, dtype=torch.int64)
    safe = torch.zeros(n, dtype=torch.uint8, device=dev); cnt = torch.zeros(n, dtype=torch.int32, device=dev)
    leaf = torch.zeros((n, 5), device=dev); cost = torch.zeros((n, 3), device=dev)
    os.environ["AUVRRT_EDGES_BRUTE"] = "1" if "allpairs" in what else "0"
    reps = int(os.environ.get("MICRO_REPS", "3"))
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if "nocost" in what:
            adev.edges_arc_dev(env, par, sd, [2.0, 0.5, 30.0, 0.5, 2.0], safe, cnt, leaf, "f32")
        else:
            adev.edges_arc_cost_dev(env, par, sd, [2.0, 0.5, 30.0, 0.5, 2.0], -4.0, safe, cnt, leaf, cost, "f32")
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    print(what, "n", n, "edges/s %.4g" % (n / min(ts)), "times", ["%.4f" % t for t in ts], "safe", float(safe.float().mean()),
          "W", float(cnt.float().mean()), "env", {k: os.environ[k] for k in os.environ if k.startswith("AUVRRT_")})
else:
    tx = torch.rand(n, device=dev) * 550 - 467; ty = torch.rand(n, device=dev) * 345 - 153
    qx = torch.tensor([-200.0], device=dev); qy = torch.tensor([0.0], device=dev)
    idx = torch.zeros(1, dtype=torch.int32, device=dev); scr = adev.nn_scratch(1, dev)
    for _ in range(3):
        adev.nn_dev(tx, ty, qx, qy, idx, scr, "f32")
    torch.cuda.synchronize()
print("done")
