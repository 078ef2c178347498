"""Drop-in for gym_rrt/envs/motion_plan_state_rrt.py (/root/reference/gym_rrt/envs/motion_plan_state_rrt.py:3-23):
the node / waypoint / obstacle record, a plain attribute bag (the boundary type callers pass in)."""


class Motion_plan_state:
    __slots__ = ("x", "y", "z", "theta", "v", "w", "traj_time_stamp", "plan_time_stamp", "size", "rl_state_id",
                 "parent", "path", "length", "cost")

    def __init__(self, x, y, z=0, theta=0, v=0, w=0, traj_time_stamp=0, plan_time_stamp=0, size=0, rl_state_id=None):
        self.x, self.y, self.z = x, y, z
        self.theta, self.v, self.w = theta, v, w
        self.traj_time_stamp, self.plan_time_stamp = traj_time_stamp, plan_time_stamp
        self.size = size
        self.rl_state_id = rl_state_id
        self.parent = None
        self.path = []
        self.length = 0
        self.cost = []

    def __repr__(self):
        return ("MPS: [x=%s, y=%s, z=%s, theta=%s, v=%s, w=%s, trag_time=%s, plan_time=%s, state_id=%s]"
                % (self.x, self.y, self.z, self.theta, self.v, self.w, self.traj_time_stamp, self.plan_time_stamp,
                   self.rl_state_id))

    __str__ = __repr__
