"""TEST INFRASTRUCTURE ONLY (oracle shim) -- geopy<2.0 `distance.vincenty(a, b).m`.

Called by /root/reference/path_planning/catalina.py:27-28.  geopy is a third-party
dependency that is absent from /root/reference and unpinned (setup.py lists only gym), so this
restates the published Vincenty inverse formula on the WGS-84 ellipsoid with geopy 1.x's
constants (major 6378.137 km, minor 6356.7523142 km, f = 1/298.257223563, stop when
|d lambda| < 1e-11, at most 20 iterations).  PARITY UNPINNED for this function: the Cartesian
map it produces is frozen as tests/golden/catalina_map.json and treated as an INPUT.
"""
import math

_A = 6378.137
_B = 6356.7523142
_F = 1 / 298.257223563


class vincenty:
    def __init__(self, p1, p2):
        self.km = self._measure(p1, p2)

    @property
    def m(self):
        return self.km * 1000.0

    @property
    def meters(self):
        return self.m

    @staticmethod
    def _measure(a, b):
        lat1, lng1 = math.radians(a[0]), math.radians(a[1])
        lat2, lng2 = math.radians(b[0]), math.radians(b[1])
        major, minor, f = _A, _B, _F
        delta_lng = lng2 - lng1
        reduced_lat1 = math.atan((1 - f) * math.tan(lat1))
        reduced_lat2 = math.atan((1 - f) * math.tan(lat2))
        sin_reduced1, cos_reduced1 = math.sin(reduced_lat1), math.cos(reduced_lat1)
        sin_reduced2, cos_reduced2 = math.sin(reduced_lat2), math.cos(reduced_lat2)
        lambda_lng = delta_lng
        lambda_prime = 2 * math.pi
        iter_limit = 20
        i = 0
        while (i == 0 or (abs(lambda_lng - lambda_prime) > 10e-12 and i <= iter_limit)):
            i += 1
            sin_lambda_lng, cos_lambda_lng = math.sin(lambda_lng), math.cos(lambda_lng)
            sin_sigma = math.sqrt(
                (cos_reduced2 * sin_lambda_lng) ** 2 +
                (cos_reduced1 * sin_reduced2 -
                 sin_reduced1 * cos_reduced2 * cos_lambda_lng) ** 2)
            if sin_sigma == 0:
                return 0.0
            cos_sigma = (sin_reduced1 * sin_reduced2 +
                         cos_reduced1 * cos_reduced2 * cos_lambda_lng)
            sigma = math.atan2(sin_sigma, cos_sigma)
            sin_alpha = cos_reduced1 * cos_reduced2 * sin_lambda_lng / sin_sigma
            cos_sq_alpha = 1 - sin_alpha ** 2
            if cos_sq_alpha != 0:
                cos2_sigma_m = cos_sigma - 2 * (sin_reduced1 * sin_reduced2 / cos_sq_alpha)
            else:
                cos2_sigma_m = 0.0
            C = f / 16. * cos_sq_alpha * (4 + f * (4 - 3 * cos_sq_alpha))
            lambda_prime = lambda_lng
            lambda_lng = (delta_lng + (1 - C) * f * sin_alpha *
                          (sigma + C * sin_sigma *
                           (cos2_sigma_m + C * cos_sigma *
                            (-1 + 2 * cos2_sigma_m ** 2))))
        u_sq = cos_sq_alpha * (major ** 2 - minor ** 2) / minor ** 2
        A = 1 + u_sq / 16384. * (4096 + u_sq * (-768 + u_sq * (320 - 175 * u_sq)))
        B = u_sq / 1024. * (256 + u_sq * (-128 + u_sq * (74 - 47 * u_sq)))
        delta_sigma = (B * sin_sigma *
                       (cos2_sigma_m + B / 4. *
                        (cos_sigma * (-1 + 2 * cos2_sigma_m ** 2) -
                         B / 6. * cos2_sigma_m * (-3 + 4 * sin_sigma ** 2) *
                         (-3 + 4 * cos2_sigma_m ** 2))))
        return minor * A * (sigma - delta_sigma)
