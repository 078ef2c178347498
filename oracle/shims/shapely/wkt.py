"""TEST INFRASTRUCTURE ONLY (oracle shim) -- inert shapely.wkt."""


def loads(*a, **k):
    raise NotImplementedError
