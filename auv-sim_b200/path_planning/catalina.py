"""Drop-in for path_planning/catalina.py of auv-sim: the Catalina map (lat/lon constants) and its
conversion to the planner's Cartesian frame.

create_cartesian / create_environs keep the reference's signatures
(/root/reference/path_planning/catalina.py:9-65).  The reference calls geopy<2.0's
`distance.vincenty`, a third-party function absent from the reference tree (unpinned); here the
Vincenty inverse formula on WGS-84 is restated so no geopy is needed.  This is input preparation on
the host, not part of the GPU hot path.
"""
import math

from motion_plan_state import Motion_plan_state

_WGS84_A = 6378137.0
_WGS84_B = 6356752.3142
_WGS84_F = 1 / 298.257223563


def vincenty_m(p, q, tol=1e-11, max_iter=20):
    """Geodesic distance in metres between (lat, lon) points p and q (degrees), Vincenty inverse."""
    phi1, lam1, phi2, lam2 = map(math.radians, (p[0], p[1], q[0], q[1]))
    a, b, f = _WGS84_A, _WGS84_B, _WGS84_F
    U1 = math.atan((1 - f) * math.tan(phi1))
    U2 = math.atan((1 - f) * math.tan(phi2))
    sU1, cU1, sU2, cU2 = math.sin(U1), math.cos(U1), math.sin(U2), math.cos(U2)
    L = lam2 - lam1
    lam = L
    for it in range(max_iter + 1):
        sl, cl = math.sin(lam), math.cos(lam)
        sin_sigma = math.hypot(cU2 * sl, cU1 * sU2 - sU1 * cU2 * cl)
        if sin_sigma == 0:
            return 0.0
        cos_sigma = sU1 * sU2 + cU1 * cU2 * cl
        sigma = math.atan2(sin_sigma, cos_sigma)
        sin_alpha = cU1 * cU2 * sl / sin_sigma
        cos2_alpha = 1 - sin_alpha * sin_alpha
        cos_2sm = cos_sigma - 2 * sU1 * sU2 / cos2_alpha if cos2_alpha != 0 else 0.0
        C = f / 16 * cos2_alpha * (4 + f * (4 - 3 * cos2_alpha))
        lam_new = L + (1 - C) * f * sin_alpha * (
            sigma + C * sin_sigma * (cos_2sm + C * cos_sigma * (2 * cos_2sm * cos_2sm - 1)))
        done = abs(lam_new - lam) <= tol
        lam = lam_new
        if done:
            break
    u2 = cos2_alpha * (a * a - b * b) / (b * b)
    A = 1 + u2 / 16384 * (4096 + u2 * (-768 + u2 * (320 - 175 * u2)))
    B = u2 / 1024 * (256 + u2 * (-128 + u2 * (74 - 47 * u2)))
    d_sigma = B * sin_sigma * (cos_2sm + B / 4 * (
        cos_sigma * (2 * cos_2sm * cos_2sm - 1)
        - B / 6 * cos_2sm * (4 * sin_sigma * sin_sigma - 3) * (4 * cos_2sm * cos_2sm - 3)))
    return b * A * (sigma - d_sigma)


def create_cartesian(pos, origin):
    """(lat, lon) -> (x, y) metres east / north of `origin` (lat, lon)."""
    lat0, lon0 = origin
    lat, lon = pos
    sx = (lon > lon0) - (lon < lon0)
    sy = (lat > lat0) - (lat < lat0)
    x = sx * vincenty_m((lat0, lon), (lat0, lon0))
    y = sy * vincenty_m((lat, lon0), (lat0, lon0))
    return (float(x), float(y))


def create_environs(obstacles, boundaries, boats, habitats):
    """lat/lon Motion_plan_state lists -> [obstacle_list, boundary_list, boat_list, habitat_list]
    in Cartesian coordinates relative to ORIGIN_BOUND."""
    def conv(items, sized):
        out = []
        for m in items:
            x, y = create_cartesian((m.x, m.y), ORIGIN_BOUND)
            out.append(Motion_plan_state(x, y, size=m.size) if sized else Motion_plan_state(x, y))
        return out
    return [conv(obstacles, True), conv(boundaries, False), conv(boats, True), conv(habitats, True)]


def _mps(rows, sized=True):
    return [Motion_plan_state(r[0], r[1], size=r[2]) if sized else Motion_plan_state(r[0], r[1]) for r in rows]


# ---- map data (lat, lon[, size in metres]); values are the Catalina survey constants of auv-sim ----
ORIGIN_BOUND = (33.445142, -118.484609)
START = (33.445170, -118.484080)
GOAL = (33.445914, -118.489636)

BOUNDARIES = _mps([(33.445914, -118.489636), (33.446866, -118.488471), (33.445064, -118.483723),
                   (33.443758, -118.485219), (33.444783, -118.488223)], sized=False)

OBSTACLES = _mps([
    (33.445113, -118.484508, 4.479407446738455), (33.445101, -118.484462, 4.337794705821955),
    (33.445088, -118.484418, 4.676061518712341), (33.445073, -118.484371, 8.238572668143059),
    (33.445047, -118.484288, 9.775751439799537), (33.445013, -118.484191, 8.625835109500015),
    (33.444986, -118.484104, 10.679512620952853), (33.444951, -118.483997, 12.150680733968928),
    (33.444914, -118.483874, 13.645304514491206), (33.444862, -118.483741, 17.812248298199293),
    (33.444779, -118.483577, 26.601714064649762)])

BOATS = _mps([
    (33.445425, -118.486314, 6), (33.444596, -118.485285, 10), (33.443940, -118.485384, 5),
    (33.444678, -118.485429, 6), (33.444820, -118.48756, 9), (33.445300, -118.485870, 10),
    (33.446200, -118.484950, 5), (33.444520, -118.484732, 5), (33.445733, -118.487100, 6),
    (33.445628, -118.487389, 7), (33.445850, -118.488100, 8), (33.445358, -118.487052, 8),
    (33.445492, -118.486839, 7), (33.445281, -118.486392, 6), (33.443872, -118.487329, 9),
    (33.445218, -118.486258, 8)])

HABITATS = _mps([
    (33.445733, -118.487789, 45), (33.446198, -118.486652, 32), (33.445400, -118.485959, 20),
    (33.445287, -118.484928, 55), (33.444457, -118.485744, 20), (33.444832, -118.485764, 30),
    (33.445534, -118.486689, 25), (33.445829, -118.485234, 30), (33.445632, -118.486744, 20),
    (33.445232, -118.485344, 20)])

GOAL_LIST = [(33.444928, -118.484448), (33.444686, -118.484716), (33.444328, -118.485606),
             (33.444811, -118.486454), (33.445491, -118.486894), (33.445491, -118.487731),
             (33.446171, -118.488010), (33.446243, -118.488697), (33.445914, -118.489636)]
