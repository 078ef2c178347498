"""TEST INFRASTRUCTURE ONLY (oracle shim) -- inert matplotlib submodule."""
from . import _Anything

_a = _Anything()


def __getattr__(name):
    return getattr(_a, name)
