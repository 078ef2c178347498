"""time the nearest-node scan (2^27 nodes, 1 query): python tools/nn_bench.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "auv-sim_b200"))
from auvrrt import device as adev
dev = torch.device("cuda", 0)
n = 1 << 27
tx = torch.rand(n, device=dev) * 550 - 467; ty = torch.rand(n, device=dev) * 345 - 153
qx = torch.tensor([-200.0], device=dev); qy = torch.tensor([0.0], device=dev)
idx = torch.zeros(1, dtype=torch.int32, device=dev); scr = adev.nn_scratch(1, dev)
for _ in range(3): adev.nn_dev(tx, ty, qx, qy, idx, scr, "f32")
torch.cuda.synchronize()
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); adev.nn_dev(tx, ty, qx, qy, idx, scr, "f32"); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e-3)
d2 = (tx - qx) ** 2 + (ty - qy) ** 2
print(os.environ.get("AUVRRT_LIB", "default").split("/")[-1], "%.1f GB/s mean, %.1f best" % (8.0 * n / np.mean(ts) / 1e9, 8.0 * n / np.min(ts) / 1e9), "ok" if int(idx.item()) == int(torch.argmin(d2).item()) else "MISMATCH")
