"""Device-resident entry points on torch tensors.

torch is plumbing here (device memory, streams, torch.distributed); the arithmetic is in the
kernels behind the `_dev` functions of include/auvrrt.h.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import F32, F64, PlanRecord, check, lib
from .api import RECORD_DTYPE, Env, _prec


def _rdtype(precision):
    return torch.float32 if _prec(precision) == F32 else torch.float64


def _stream_ptr(stream=None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def _vp(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class DevicePlanner:
    """Batched RRT.exploring with everything resident in HBM: starts/seeds in, 96-byte records
    (+ chains of stream positions) out.  One instance owns the tree workspace."""

    def __init__(self, env: Env, params, precision="f32", max_queries=4096, want_chain=True, want_path=False):
        self.env, self.params, self.prec = env, params, _prec(precision)
        self.dev = torch.device("cuda", env.device)
        self.rdtype = _rdtype(precision)
        wsb = lib().auvrrt_plan_workspace_bytes_q(env.handle, C.byref(params), self.prec, int(max_queries))
        if wsb < 0:
            raise _lib.AuvrrtError(lib().auvrrt_last_error().decode())
        self.workspace = torch.empty(int(wsb), dtype=torch.uint8, device=self.dev)
        self.Q = int(max_queries)
        self.starts = torch.zeros((self.Q, 5), dtype=self.rdtype, device=self.dev)
        self.seeds = torch.zeros(self.Q, dtype=torch.int64, device=self.dev)      # uint64 bit patterns
        self.records = torch.zeros((self.Q, C.sizeof(PlanRecord)), dtype=torch.uint8, device=self.dev)
        ccap = max(params.chain_cap, 1)
        self.chain = torch.zeros((self.Q, ccap), dtype=torch.int32, device=self.dev) if (want_chain or want_path) else None
        self.path = (torch.zeros((self.Q, params.path_cap, 6), dtype=self.rdtype, device=self.dev)
                     if want_path and params.path_cap > 0 else None)

    def set_queries(self, starts: np.ndarray, seeds: np.ndarray):
        q = len(seeds)
        assert q <= self.Q
        self.starts[:q].copy_(torch.from_numpy(np.asarray(starts, dtype=np.float64)).to(self.rdtype))
        self.seeds[:q].copy_(torch.from_numpy(np.asarray(seeds, dtype=np.uint64).view(np.int64)))
        self.n = q

    def launch(self, stream=None):
        """enqueue one pass over the current queries on `stream`; no synchronisation"""
        check(lib().auvrrt_plan_batch_dev(self.env.handle, _vp(self.starts), _vp(self.seeds), self.n,
                                          C.byref(self.params), self.prec, _vp(self.workspace),
                                          self.workspace.numel(), _vp(self.records), _vp(self.chain),
                                          _vp(self.path), None, _stream_ptr(stream)))

    def records_numpy(self):
        return self.records[:self.n].cpu().numpy().view(RECORD_DTYPE).reshape(-1)


def nn_dev(tree_x, tree_y, qx, qy, out_idx, scratch, precision, stream=None):
    check(lib().auvrrt_nn_dev(_vp(tree_x), _vp(tree_y), tree_x.numel(), _vp(qx), _vp(qy), qx.numel(),
                              _prec(precision), _vp(scratch), scratch.numel(), _vp(out_idx), _stream_ptr(stream)))


def nn_scratch(nq, device):
    return torch.empty(int(lib().auvrrt_nn_scratch_bytes(int(nq))), dtype=torch.uint8, device=device)


def edges_dubins_dev(env: Env, q0, q1, rho, W, out_safe, out_word, out_length, precision, stream=None):
    check(lib().auvrrt_edges_dubins_dev(env.handle, _vp(q0), _vp(q1), q0.shape[0], float(rho), int(W),
                                        _prec(precision), _vp(out_safe), _vp(out_word), _vp(out_length),
                                        _stream_ptr(stream)))


def edges_dubins_cost_dev(env: Env, q0, q1, rho, W, velocity, w3, out_safe, out_word, out_length, out_cost, precision, stream=None):
    check(lib().auvrrt_edges_dubins_cost_dev(env.handle, _vp(q0), _vp(q1), q0.shape[0], float(rho), int(W), float(velocity),
                                             float(w3), _prec(precision), _vp(out_safe), _vp(out_word), _vp(out_length),
                                             _vp(out_cost), _stream_ptr(stream)))


def edges_arc_cost_dev(env: Env, parents, seeds, params5, w3, out_safe, out_counts, out_leaf, out_cost, precision, stream=None):
    p = (C.c_double * 5)(*[float(x) for x in params5])
    check(lib().auvrrt_edges_arc_cost_dev(env.handle, _vp(parents), _vp(seeds), parents.shape[0], p, float(w3),
                                          _prec(precision), _vp(out_safe), _vp(out_counts), _vp(out_leaf), _vp(out_cost),
                                          _stream_ptr(stream)))


def edges_arc_dev(env: Env, parents, seeds, params5, out_safe, out_counts, out_leaf, precision, stream=None):
    p = (C.c_double * 5)(*[float(x) for x in params5])
    check(lib().auvrrt_edges_arc_dev(env.handle, _vp(parents), _vp(seeds), parents.shape[0], p,
                                     _prec(precision), _vp(out_safe), _vp(out_counts), _vp(out_leaf),
                                     _stream_ptr(stream)))
