"""TEST INFRASTRUCTURE ONLY (oracle shim) -- inert shapely.ops."""


def split(*a, **k):
    raise NotImplementedError("shapely.ops.split is not on the RRT hot path; shim is inert")
