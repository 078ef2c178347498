"""auvrrt -- B200 (sm_100a) implementation of the RRT planning inner loop of auv-sim.

`auvrrt.api` wraps the host-buffer C ABI (numpy in, numpy out); `auvrrt.device` the device-pointer
entry points on torch tensors (plumbing only: memory, streams, torch.distributed);
`path_planning/` next to this package holds the drop-in `rrt_dubins` / `cost` modules.
"""
from ._lib import F32, F64, AuvrrtError, LIB_PATH, lib  # noqa: F401
from . import api  # noqa: F401
from .api import Env, plan_params  # noqa: F401
