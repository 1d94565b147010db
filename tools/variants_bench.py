#!/usr/bin/env python
"""K1 throughput of the non-headline BASELINE configs: CartPole POMDP + GRU (P = 4097) and simple_spread
(P = 16384, N = 2 / 3), plus the 2^20 CartPole genetic population.  Prints one JSON line per case."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_es_b200.engine import RolloutEngine  # noqa: E402


def timed(eng, gen0, sigma, parents, reps=3):
    P = eng.P
    fit = torch.zeros(P, dtype=torch.float64, device="cuda"); steps = torch.zeros(P, dtype=torch.int64, device="cuda")
    eng.rollout(gen0, sigma, parents, fitness=fit, steps=steps)
    torch.cuda.synchronize()
    best = None
    for r in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); eng.rollout(gen0 + 1 + r, sigma, parents, fitness=fit, steps=steps); e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1); n = int(steps.sum().item())
        if best is None or ms < best[0]:
            best = (ms, n, float(fit.max().item()), float(fit.mean().item()))
    return {"ms": best[0], "env_steps": best[1], "steps_per_s": best[1] / (best[0] * 1e-3), "best": best[2], "mean": best[3]}


def gru_balancing_parent():
    """A GRU policy that balances the (fully observed) pole: fc1 unit 0 = bang-bang feature, the n gate carries it,
    z ~ 0 -- every episode runs the full 500 steps (the converged regime of the GRU rollout kernel)."""
    mu = np.zeros(6562, np.float32)
    W1 = mu[0:128].reshape(32, 4)
    Wih = mu[160:160 + 3072].reshape(96, 32)
    bih = mu[160 + 6144:160 + 6144 + 96]
    W2 = mu[160 + 6144 + 192:160 + 6144 + 192 + 64].reshape(2, 32)
    W1[0] = [0.0, 0.5, 10.0, 3.0]
    Wih[64, 0] = 3.0
    bih[32:64] = -10.0
    W2[1, 0] = 5.0; W2[0, 0] = -5.0
    return mu[None]


def main():
    out = {}
    rng = np.random.default_rng(0)
    if len(sys.argv) > 1 and sys.argv[1] == "gru_converged":
        eng = RolloutEngine("CartPole-v1", 4, 2, True, False, 500, 5, 4097, 4097, 2, 1, seed=0)
        print(json.dumps({"gru_converged": timed(eng, 0, 0.02, torch.from_numpy(gru_balancing_parent()).cuda(), reps=3)}))
        eng.close()
        eng = RolloutEngine("CartPole-v1", 4, 2, True, True, 500, 5, 4097, 4097, 2, 1, seed=0)
        print(json.dumps({"gru_gen0": timed(eng, 0, 1.0, torch.zeros(1, 6562, dtype=torch.float32, device="cuda"), reps=3)}))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "spread3":
        eng = RolloutEngine("simple_spread", 18, 5, False, False, "None", 5, 16384, 16384, 1, 1, seed=0, n_agents=3, init_mode="fresh")
        print(json.dumps({"spread_n3": timed(eng, 0, 0.2, torch.zeros(1, 6 * 3 * 32 + 32 + 165, dtype=torch.float32, device="cuda"), reps=2)}))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "classic":
        for env, obs, act in (("MountainCar-v0", 2, 3), ("Acrobot-v1", 6, 3)):
            eng = RolloutEngine(env, obs, act, False, False, None, 5, 16384, 16384, 1, 1, seed=0, init_mode="fresh")
            print(json.dumps({env: timed(eng, 0, 2.0, torch.zeros(1, eng.D, dtype=torch.float32, device="cuda"), reps=2)}))
        return
    eng = RolloutEngine("CartPole-v1", 4, 2, True, False, 500, 5, 4097, 4097, 2, 1, seed=0)
    out["gru_converged"] = timed(eng, 0, 0.02, torch.from_numpy(gru_balancing_parent()).cuda())
    eng.close()
    # config 2: CartPole POMDP GRU, simple_evolution, P = 4097
    for label, scale, sigma in (("gru_gen0", 0.0, 1.0), ("gru_random_parent", 0.3, 0.2)):
        eng = RolloutEngine("CartPole-v1", 4, 2, True, True, 500, 5, 4097, 4097, 2, 1, seed=0)
        mu = torch.from_numpy((rng.normal(0, 1, (1, 6562)) * scale).astype(np.float32)).cuda()
        out[label] = timed(eng, 0, sigma, mu)
        eng.close()
    # config 4: simple_spread openai_es P = 16384
    for N in (2, 3):
        D = 6 * N * 32 + 32 + 165
        eng = RolloutEngine("simple_spread", 6 * N, 5, False, False, "None", 5, 16384, 16384, 1, 1, seed=0, n_agents=N, init_mode="fresh")
        mu = torch.zeros(1, D, dtype=torch.float32, device="cuda")
        out["spread_n%d" % N] = timed(eng, 0, 0.2, mu)
        eng.close()
    # classic control beyond CartPole (conf/mountaincar.yaml, conf/acrobot.yaml): P = 16384, random policies
    for env, obs, act in (("MountainCar-v0", 2, 3), ("Acrobot-v1", 6, 3)):
        eng = RolloutEngine(env, obs, act, False, False, None, 5, 16384, 16384, 1, 1, seed=0, init_mode="fresh")
        mu = torch.zeros(1, eng.D, dtype=torch.float32, device="cuda")
        out[env] = timed(eng, 0, 2.0, mu)
        eng.close()
    # config 5: CartPole simple_genetic P = 2^20 (16 elites), gen-0 regime
    eng = RolloutEngine("CartPole-v1", 4, 2, False, False, 500, 5, 1 << 20, (1 << 20) // 16, 1, 16, seed=0)
    par = torch.zeros(16, 226, dtype=torch.float32, device="cuda")
    out["genetic_2^20_gen0"] = timed(eng, 0, 1.0, par, reps=2)
    eng.close()
    for k, v in out.items():
        print(json.dumps({k: v}))


if __name__ == "__main__":
    main()
