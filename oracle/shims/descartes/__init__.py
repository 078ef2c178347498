"""TEST INFRASTRUCTURE ONLY (oracle shim) -- inert descartes."""


class PolygonPatch:
    def __init__(self, *a, **k):
        pass
