"""Vectorised gym_rrt Planner_RRT on the GPU (libauvrrt.so, csrc/gym.cu): Q independent episodes of
gym_rrt/envs/rrt_dubins.py:Planner_RRT resident in HBM, stepped together.

`GymBatch.plan(max_step)` is Planner_RRT.planning for every episode; `GymBatch.step(actions)` is what
RRTEnv.step does with the agent's chosen sub-cell (gym_rrt/envs/rrt_env.py:206-247).  No CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import F32, F64, GymParams, GymRecord, check, lib

GYM_RECORD_DTYPE = np.dtype([
    ("status", "<i4"), ("done", "<i4"), ("steps", "<i4"), ("n_nodes", "<i4"), ("n_occupied", "<i4"),
    ("last_parent", "<i4"), ("last_accepted", "<i4"), ("last_nwp", "<i4"), ("last_uniforms", "<i4"),
    ("n_path", "<i4"), ("n_uniforms", "<i8"), ("cand", "<f8", (4,)), ("arc_length", "<f8")])
assert GYM_RECORD_DTYPE.itemsize == C.sizeof(GymRecord) == 88

# rewards of RRTEnv.step (gym_rrt/envs/rrt_env.py:40-42, :236-245)
R_FOUND_PATH, R_CREATE_NODE, R_INVALID_NODE = 300, 0, -1


def _f64(a, shape):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(shape))


class GymBatch:
    def __init__(self, boundary, obstacles, n_episodes, *, exp_rate=1.0, dist_to_end=2.0, diff_max=0.5, freq=50.0,
                 cell_side_length=2.0, subsections_in_cell=8, node_cap=201, track_counts=False, precision=F32, device=0):
        """boundary = (x0, y0, x1, y1); obstacles = [(x, y, size)] in obstacle_list order."""
        circles = _f64(obstacles, (-1, 3)) if len(obstacles) else np.zeros((0, 3))
        self.params = GymParams(float(boundary[0]), float(boundary[1]), float(boundary[2]), float(boundary[3]),
                                float(exp_rate), float(dist_to_end), float(diff_max), float(freq),
                                float(cell_side_length), int(subsections_in_cell), int(node_cap), int(bool(track_counts)), 0)
        self.Q = int(n_episodes)
        self.precision = precision
        self.device = device
        self._h = C.c_void_p()
        check(lib().auvrrt_gym_create(circles.ctypes.data_as(_lib._dp), len(circles), C.byref(self.params),
                                      C.c_int64(self.Q), int(precision), int(device), C.byref(self._h)))
        r, c = C.c_int(), C.c_int()
        check(lib().auvrrt_gym_grid_shape(self._h, C.byref(r), C.byref(c)))
        self.rows, self.cols, self.subsections = r.value, c.value, int(subsections_in_cell)
        self.n_subcells = self.rows * self.cols * self.subsections
        self.node_cap = int(node_cap)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().auvrrt_gym_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def flat_id(self, row, col, sub):
        return (int(row) * self.cols + int(col)) * self.subsections + int(sub)

    def reset(self, starts, goals, seeds):
        starts, goals = _f64(starts, (self.Q, 3)), _f64(goals, (self.Q, 2))
        seeds = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64).reshape(self.Q))
        check(lib().auvrrt_gym_reset(self._h, starts.ctypes.data_as(_lib._dp), goals.ctypes.data_as(_lib._dp),
                                     seeds.ctypes.data_as(_lib._u64p)))

    def _run(self, actions, n_steps, full_candidates, want_records):
        recs = np.zeros(self.Q, GYM_RECORD_DTYPE) if want_records else None
        ap = None
        if actions is not None:
            actions = np.ascontiguousarray(np.asarray(actions, dtype=np.int32).reshape(self.Q))
            ap = actions.ctypes.data_as(_lib._i32p)
        check(lib().auvrrt_gym_step(self._h, ap, int(n_steps), int(bool(full_candidates)),
                                    recs.ctypes.data_as(C.c_void_p) if want_records else None))
        return recs

    def plan(self, max_step=200, *, full_candidates=False, records=True):
        """Planner_RRT.planning(max_step) for every episode (cells drawn with random.choice)."""
        return self._run(None, max_step, full_candidates, records)

    def step(self, actions, *, full_candidates=False, records=True):
        """RRTEnv.step: actions[Q] are flat sub-cell ids ((row * cols + col) * subsections + sub)."""
        return self._run(actions, 1, full_candidates, records)

    @staticmethod
    def rewards(recs):
        """RRTEnv.step's reward (rrt_env.py:236-245) from the records of one step() call."""
        return np.where(recs["done"] != 0, R_FOUND_PATH,
                        np.where(recs["last_accepted"] != 0, R_CREATE_NODE, R_INVALID_NODE)).astype(np.int32)

    def tree(self, q):
        nodes = np.zeros((self.node_cap, 4))
        parents = np.zeros(self.node_cap, np.int32)
        cells = np.zeros(self.node_cap, np.int32)
        occupied = np.zeros(self.node_cap, np.int32)
        n, no = C.c_int32(), C.c_int32()
        check(lib().auvrrt_gym_tree(self._h, C.c_int64(q), self.node_cap, nodes.ctypes.data_as(_lib._dp),
                                    parents.ctypes.data_as(_lib._i32p), cells.ctypes.data_as(_lib._i32p),
                                    occupied.ctypes.data_as(_lib._i32p), C.byref(n), C.byref(no)))
        return {"nodes": nodes[:n.value].copy(), "parents": parents[:n.value].copy(), "cells": cells[:n.value].copy(),
                "occupied": occupied[:no.value].copy()}

    def counts(self, q0=0, nq=None):
        """node count of every sub-cell: RRTEnv.convert_rrt_grid_to_1D_num_of_nodes_only (rrt_env.py:265-277)"""
        nq = self.Q - q0 if nq is None else nq
        out = np.zeros((nq, self.n_subcells), np.uint16)
        check(lib().auvrrt_gym_counts(self._h, C.c_int64(q0), C.c_int64(nq), out.ctypes.data_as(C.POINTER(C.c_uint16))))
        return out

    def counts_device_ptr(self):
        return lib().auvrrt_gym_counts_dev(self._h)

    def path(self, q, cap=8192):
        """generate_final_course of a done episode: [n][3] x, y, theta, goal arc end -> start."""
        out = np.zeros((cap, 3))
        n = C.c_int32()
        check(lib().auvrrt_gym_path(self._h, C.c_int64(q), cap, out.ctypes.data_as(_lib._dp), C.byref(n)))
        return out[:n.value].copy()
