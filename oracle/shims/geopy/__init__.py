"""TEST INFRASTRUCTURE ONLY (oracle shim for geopy<2.0) -- see distance.py."""
from . import distance  # noqa: F401
