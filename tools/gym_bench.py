"""Quick throughput look at the vectorised gym planner (not the bench contract; see bench.py)."""
import sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "auv-sim_b200")); sys.path.insert(0, ROOT)
from auvrrt import gym
from oracle.harness import GYM_MAIN_OBSTACLES

def episodes(Q, seed):
    rs = np.random.default_rng(seed)
    starts = np.column_stack([rs.uniform(5, 15, Q), rs.uniform(5, 15, Q), rs.uniform(-np.pi, np.pi, Q)])
    goals = np.column_stack([rs.uniform(35, 45, Q), rs.uniform(35, 45, Q)])
    return starts, goals, np.arange(Q, dtype=np.uint64)

for prec, name in ((gym.F32, "f32"), (gym.F64, "f64")):
    for Q in (4096, 65536, 262144):
        for freq in (10.0, 50.0):
            if prec == gym.F64 and Q > 65536: continue
            b = gym.GymBatch((0, 0, 50, 50), GYM_MAIN_OBSTACLES, Q, freq=freq, node_cap=201, precision=prec)
            s, g, sd = episodes(Q, 1)
            best = 1e9
            for rep in range(3):
                b.reset(s, g, sd)
                t0 = time.perf_counter(); r = b.plan(200, records=True); dt = time.perf_counter() - t0
                best = min(best, dt)
            steps = int(r["steps"].sum())
            print(f"{name} Q={Q} freq={freq}: {best*1e3:.2f} ms, steps={steps} ({steps/best:.3e} steps/s), done={r['done'].mean():.3f}, uniforms={int(r['n_uniforms'].sum())/best:.3e}/s", flush=True)
            b.close()
