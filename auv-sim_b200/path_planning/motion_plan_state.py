"""Drop-in for path_planning/motion_plan_state.py of auv-sim: the node / waypoint / obstacle record.

Same constructor and attributes as the reference class (/root/reference/path_planning/
motion_plan_state.py:6-19); it stays a plain Python attribute bag because that is the boundary
type callers hand to RRT and the cost functions.  The arithmetic lives in libauvrrt.so.
"""


class Motion_plan_state:
    __slots__ = ("x", "y", "z", "theta", "v", "w", "traj_time_stamp", "plan_time_stamp", "size",
                 "parent", "path", "length", "cost")

    def __init__(self, x, y, z=0, theta=0, v=0, w=0, traj_time_stamp=0, plan_time_stamp=0, size=0, length=0):
        self.x, self.y, self.z = x, y, z
        self.theta, self.v, self.w = theta, v, w
        self.traj_time_stamp, self.plan_time_stamp = traj_time_stamp, plan_time_stamp
        self.size = size
        self.parent = None
        self.path = []
        self.length = length
        self.cost = []

    def _kind(self):
        stamps = self.traj_time_stamp == 0 and self.plan_time_stamp == 0
        still = self.theta == 0 and self.v == 0 and self.w == 0
        if self.z == 0 and still and stamps:
            return "xy"
        if still and self.size == 0 and stamps:
            return "xyz"
        if self.size != 0 and stamps:
            return "obstacle"
        if self.z == 0 and self.v == 0 and self.w == 0:
            return "pose"
        return "full"

    def __repr__(self):
        k = self._kind()
        if k == "xy":
            body = "x=%s, y=%s" % (self.x, self.y)
        elif k == "xyz":
            body = "x=%s, y=%s, z=%s" % (self.x, self.y, self.z)
        elif k == "obstacle":
            body = "x=%s, y=%s, z=%s, size=%s" % (self.x, self.y, self.z, self.size)
        elif k == "pose":
            body = "x=%s, y=%s, theta=%s, trag_time=%s, plan_time=%s" % (
                self.x, self.y, self.theta, self.traj_time_stamp, self.plan_time_stamp)
        else:
            body = "x=%s, y=%s, z=%s, theta=%s, v=%s, w=%s, trag_time=%s, plan_time=%s" % (
                self.x, self.y, self.z, self.theta, self.v, self.w, self.traj_time_stamp, self.plan_time_stamp)
        return "MPS: [" + body + "]"

    __str__ = __repr__
