#!/usr/bin/env python
"""Small rollouts of every K1 kernel + K2 + K3 for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_es_b200.engine import RolloutEngine  # noqa: E402

CASES = [("CartPole-v1", 4, 2, False, 5, 1), ("CartPole-v1", 4, 2, False, 3, 1), ("CartPole-v1", 4, 2, True, 5, 1),
         ("simple_spread", 12, 5, False, 5, 2), ("simple_spread", 18, 5, False, 2, 3), ("MountainCar-v0", 2, 3, False, 5, 1),
         ("Acrobot-v1", 6, 3, False, 4, 1)]
for env, obs, act, gru, E, N in CASES:
    P = 200
    eng = RolloutEngine(env, obs, act, gru, False, None, E, P, P, 1, 1, seed=1, n_agents=N, init_mode="fresh")
    mu = torch.zeros(1, eng.D, dtype=torch.float32, device="cuda")
    fit, steps, trace, actions = eng.rollout(0, 1.0, mu, n_trace=4)
    order, shaped = eng.rank_desc(fit, shaped=True, full_key=True)
    m = torch.zeros(eng.D, dtype=torch.float32, device="cuda"); v = torch.zeros_like(m)
    eng.update_openai(0, 1.0, 0.1, 1, shaped, mu.view(-1), m, v)
    eng.elite_mean(0, 1.0, mu, order, 5)
    eng.materialize(0, 1.0, mu, order[:7].contiguous())
    torch.cuda.synchronize()
    print(env, "gru" if gru else "mlp", "E", E, "steps", int(steps.sum()), "best", float(fit.max()))
    eng.close()
print("sanitize run complete")
