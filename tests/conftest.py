"""pytest configuration: `gpu` marker and import paths.

`-m "not gpu"` runs on the CPU-only build container; `-m gpu` needs a B200 and calls the CUDA
path through the C ABI (libauvrrt.so).
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "auv-sim_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 GPU (run with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def catalina_map():
    with open(os.path.join(GOLDEN, "catalina_map.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def shark_grid():
    z = np.load(os.path.join(GOLDEN, "shark_grid.npz"))
    return z["bins"], z["probs"]


@pytest.fixture(scope="session")
def exploring_golden():
    z = np.load(os.path.join(GOLDEN, "exploring.npz"))
    meta = json.loads(str(z["meta"]))
    return z, meta


class GymEpisode:
    """One golden episode of gym_rrt Planner_RRT (tests/golden/gym_plan.npz, oracle/make_golden_gym.py)."""

    def __init__(self, z, name):
        s = z[name + "/setup"]
        self.name = name
        self.start, self.goal, self.boundary = s[0:3].copy(), s[3:5].copy(), s[5:9].copy()
        self.freq, self.cell_side, self.subsections = float(s[9]), float(s[10]), int(s[11])
        self.max_step, self.seed, self.steps, self.found = int(s[12]), int(s[13]), int(s[14]), bool(s[15])
        self.goal_arc_length = float(s[16])
        self.obstacles = z[name + "/obstacles"]
        self.rows, self.cols = [int(v) for v in z[name + "/grid_shape"]]
        self.counts_nz = z[name + "/counts_nz"]
        self.actions = z[name + "/actions"] if (name + "/actions") in z.files else None
        for k in ("parent", "nwp", "accepted", "done", "n_nodes", "n_occupied", "n_uniforms", "cand", "occupied",
                  "nodes", "path"):
            setattr(self, k, z[name + "/" + k])

    def flat(self, rck):
        rck = np.asarray(rck).reshape(-1, 3)
        rck = np.column_stack([rck[:, 0] % self.rows, rck[:, 1] % self.cols, rck[:, 2]])   # list[-k] indexing
        return ((rck[:, 0] * self.cols + rck[:, 1]) * self.subsections + rck[:, 2]).astype(np.int32)

    def flat_actions(self):
        """actions padded to max_step with -1 (an id no cell has: skipped like an empty cell)"""
        a = np.full(self.max_step, -1, np.int32)
        f = self.flat(self.actions)
        a[:len(f)] = f
        return a


@pytest.fixture(scope="session")
def gym_golden():
    z = np.load(os.path.join(GOLDEN, "gym_plan.npz"))
    return [GymEpisode(z, str(n)) for n in z["names"]]


class AstarCase:
    """One golden query of astar_fixLenSOG.astar (tests/golden/astar.npz, oracle/make_golden_astar.py)."""

    def __init__(self, z, name):
        s = z[name + "/setup"]
        self.name = name
        self.start, self.limit, self.weights = s[0:2].copy(), float(s[2]), s[3:7].copy()
        self.velocity, self.n_bins = float(s[7]), int(s[8])
        self.outcome = str(z[name + "/outcome"])
        self.expanded = z[name + "/expanded"]
        self.n_visited = int(z[name + "/n_visited"])
        if self.outcome == "ok":
            self.nodes, self.cost_list = z[name + "/nodes"], z[name + "/cost_list"]
            self.smooth_keep, self.smooth_path = z[name + "/smooth_keep"], z[name + "/smooth_path"]
            self.cost = float(z[name + "/cost"])

    @property
    def want_status(self):
        return {"ok": 0, "none": 1}.get(self.outcome, 3)


@pytest.fixture(scope="session")
def astar_golden(catalina_map, shark_grid):
    z = np.load(os.path.join(GOLDEN, "astar.npz"))
    bins, probs = shark_grid
    world = dict(circles=catalina_map["circles"], boundary=catalina_map["boundary"], habitats=catalina_map["habitats"],
                 centroid=z["centroid"], cells_rounded=z["cells_rounded"], bins=bins, probs=probs)
    return world, [AstarCase(z, str(n)) for n in z["names"]]
