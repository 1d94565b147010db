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

# (env, obs, act, gru, E, n_agents, continuous)
CASES = [("CartPole-v1", 4, 2, False, 5, 1, False), ("CartPole-v1", 4, 2, False, 3, 1, False), ("CartPole-v1", 4, 2, True, 5, 1, False),
         ("simple_spread", 12, 5, False, 5, 2, False), ("simple_spread", 18, 5, False, 2, 3, False), ("MountainCar-v0", 2, 3, False, 5, 1, False),
         ("Acrobot-v1", 6, 3, False, 4, 1, False),
         # round 2: continuous head, the generic GRU kernel on every env
         ("Pendulum-v0", 3, 1, False, 5, 1, True), ("Pendulum-v0", 3, 1, True, 3, 1, True), ("MountainCar-v0", 2, 3, True, 5, 1, False),
         ("Acrobot-v1", 6, 3, True, 2, 1, False), ("simple_spread", 12, 5, True, 4, 2, False), ("simple_spread", 18, 5, True, 3, 3, False)]
for env, obs, act, gru, E, N, cont in CASES:
    P = 200
    eng = RolloutEngine(env, obs, act, gru, False, None, E, P, P, 1, 1, seed=1, n_agents=N, init_mode="fresh", discrete_action=not cont)
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
# round 2: the CartPole-MLP scheduler paths -- episode-granular queues (aligned queue A, exact queue B with re-alignment: the second
# launch of a case takes its tail length from the first one's mean episode length), the straggler phase (ragged generation-0
# episodes and a converged population whose last warps run sparse), sparse warps (a launch that does not fill the SMs, forced
# shares), integer-key K2
w1 = np.zeros((1, 226), np.float32)
w1[0, :4] = [0.0, 0.5, 10.0, 3.0]; w1[0, 160 + 32] = 5.0; w1[0, 160] = -5.0
KNOBS = ("SES_ROLLOUT_LANES", "SES_K1_SPARSE_RANK", "SES_K1_SPARSE_QUOTA", "SES_K1_TAIL")
for P, sigma, parent, knobs in [(3000, 2.0, np.zeros((1, 226), np.float32), {}), (9000, 0.05, w1, {}), (9000, 2.0, w1, {"SES_K1_TAIL": "1"}),
                                (8192, 0.05, w1, {}), (600, 0.05, w1, {"SES_K1_SPARSE_RANK": "0", "SES_K1_SPARSE_QUOTA": "1"}),
                                (2500, 2.0, w1, {"SES_K1_SPARSE_RANK": "1", "SES_K1_SPARSE_QUOTA": "3"})]:
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update(knobs)
    eng = RolloutEngine("CartPole-v1", 4, 2, False, False, 500, 5, P, P, 1, 1, seed=3)
    for gen in (1, 2):
        fit, steps = eng.rollout(gen, sigma, torch.from_numpy(parent).cuda())
        order, shaped = eng.rank_desc(fit, shaped=True)
        torch.cuda.synchronize()
    print("CartPole-v1 scheduler case P", P, knobs, "steps", int(steps.sum()), "best", float(fit.max()))
    eng.close()
for k in KNOBS:
    os.environ.pop(k, None)
# the test build's alternative GRU kernels (2: a warp pair per offspring over named barriers)
for variant in ("0", "1", "2", "4"):
    os.environ["SES_GRU_VARIANT"] = variant
    eng = RolloutEngine("CartPole-v1", 4, 2, True, False, None, 5, 120, 120, 1, 1, seed=1, test_build=True)
    fit, steps = eng.rollout(0, 1.0, torch.zeros(1, eng.D, dtype=torch.float32, device="cuda"))
    torch.cuda.synchronize()
    print("CartPole-v1 gru variant", variant, "steps", int(steps.sum()))
    eng.close()
os.environ.pop("SES_GRU_VARIANT", None)
print("sanitize run complete")
