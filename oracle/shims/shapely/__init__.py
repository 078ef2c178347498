"""TEST INFRASTRUCTURE ONLY (oracle shim) -- not product code.

Minimal stand-in for the `shapely` package so the UNMODIFIED reference modules
(/root/reference/path_planning/rrt_dubins.py:9, sharkOccupancyGrid.py:2-3) import in a
container without shapely.  Only what the RRT hot path touches is functional:
`Polygon(pts).bounds`, `.exterior.xy`, and `Point(x, y).within(poly)` (strict interior,
decided EXACTLY with rational arithmetic when the float filter is inconclusive, which is
what GEOS's robust orientation predicates guarantee).  Everything else is inert.
"""
