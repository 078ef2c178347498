from gym_rrt.envs.rrt_env import RRTEnv, VecRRTEnv  # noqa: F401
