#!/usr/bin/env python
"""K1 micro-benchmark: the rollout kernel alone in the two policy regimes of SURVEY.md section 8d.

  gen0      : mu = 0, sigma = 2  (mean return ~ 20 steps, ragged episode lengths)
  converged : a parent that balances the pole, sigma = 0.05 (nearly every episode runs 500 steps)

Knobs come from the environment (SES_K1_VARIANT, SES_ROLLOUT_LANES, SES_ROLLOUT_CTAS_PER_SM).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_es_b200.engine import RolloutEngine  # noqa: E402

D = 226


def balancing_parent():
    mu = np.zeros((1, D), np.float32)
    w1 = mu[0, :128].reshape(32, 4); w2 = mu[0, 160:224].reshape(2, 32)
    w1[0] = [0.0, 0.5, 10.0, 3.0]
    w2[1, 0] = 5.0; w2[0, 0] = -5.0
    return mu


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pop", type=int, default=65536)
    ap.add_argument("--E", type=int, default=5)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--regime", default="both")
    args = ap.parse_args()
    P = args.pop
    out = {"variant": os.environ.get("SES_K1_VARIANT", "7"), "lanes": os.environ.get("SES_ROLLOUT_LANES", "auto"), "ctas_per_sm": os.environ.get("SES_ROLLOUT_CTAS_PER_SM", "auto"), "pop": P}
    for regime in (["gen0", "converged"] if args.regime == "both" else [args.regime]):
        eng = RolloutEngine("CartPole-v1", 4, 2, False, False, 500, args.E, P, P, 1, 1, seed=0)
        mu = torch.from_numpy(np.zeros((1, D), np.float32) if regime == "gen0" else balancing_parent()).cuda()
        sigma = 2.0 if regime == "gen0" else 0.05
        fit = torch.zeros(P, dtype=torch.float64, device="cuda"); steps = torch.zeros(P, dtype=torch.int64, device="cuda")
        for g in range(2):
            eng.rollout(g, sigma, mu, fitness=fit, steps=steps)
        torch.cuda.synchronize()
        ts, ns = [], []
        for g in range(args.reps):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); eng.rollout(10 + g, sigma, mu, fitness=fit, steps=steps); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1)); ns.append(int(steps.sum().item()))
        i = int(np.argmin(ts))
        out[regime] = {"ms": ts[i], "env_steps": ns[i], "steps_per_s": ns[i] / (ts[i] * 1e-3), "mean_len": ns[i] / (P * args.E),
                       "ms_all": [round(t, 3) for t in ts]}
        eng.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
