"""Drop-in for auv-sim's gym_rrt package: the goal-directed RRT planner the RL environment drives,
backed by libauvrrt.so (no CPU fallback).  Unlike the reference's __init__ this one does not import
`gym` or register 'rrt-v0' unless gym is installed (the planner itself never needs it)."""
try:  # pragma: no cover - gym is optional
    from gym.envs.registration import register
    register(id='rrt-v0', entry_point='gym_rrt.envs:RRTEnv')
except Exception:  # gym absent (or the id already registered)
    pass
