#!/usr/bin/env python
"""K1 launch-geometry sweep: converged CartPole-MLP populations (every episode 500 steps) timed over
(population, lanes per warp, CTAs per SM, variant), to fit the 'rounds x per-round time' model of
DESIGN.md section 5.1.  One JSON line per point.

  python tools/k1_rounds.py --points "65536:30:3:7,65536:32:3:7,63936:30:3:7"     # P:lanes:ctas_per_sm:variant
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_es_b200.engine import RolloutEngine  # noqa: E402
from tools.k1_bench import balancing_parent, D  # noqa: E402


def run_point(P, lanes, ctas, variant, E, regime, reps, extra_env=None):
    env = {"SES_ROLLOUT_LANES": lanes, "SES_ROLLOUT_CTAS_PER_SM": ctas, "SES_K1_VARIANT": variant}
    env.update(extra_env or {})
    for k, v in env.items():
        if v in (0, "0", None, ""):
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)
    eng = RolloutEngine("CartPole-v1", 4, 2, False, False, 500, E, P, P, 1, 1, seed=0)
    mu = torch.from_numpy(np.zeros((1, D), np.float32) if regime == "gen0" else balancing_parent()).cuda()
    sigma = 2.0 if regime == "gen0" else 0.05
    fit = torch.zeros(P, dtype=torch.float64, device="cuda"); steps = torch.zeros(P, dtype=torch.int64, device="cuda")
    # keep the GPU busy (and its clocks up) before and during the measurement: back-to-back launches, one event pair around all
    busy = max(3, int(30.0 / max(0.3, P * E * 1e-5)))            # ~30 ms of warm-up launches
    for g in range(busy):
        eng.rollout(g, sigma, mu, fitness=fit, steps=steps)
    ts = []
    for r in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        inner = max(1, busy // 3)
        e0.record()
        for g in range(inner):
            eng.rollout(1000 + r * inner + g, sigma, mu, fitness=fit, steps=steps)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / inner)
    n = int(steps.sum().item())
    eng.close()
    t = min(ts)
    return {"P": P, "E": E, "lanes": lanes, "ctas_per_sm": ctas, "variant": variant, "regime": regime, "ms": round(t, 4),
            "G_steps_s": round(n / t / 1e6, 3), "env_steps": n, "extra": extra_env or {}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", required=True)
    ap.add_argument("--E", type=int, default=5)
    ap.add_argument("--regime", default="converged")
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--env", default="", help="extra KEY=VAL,KEY=VAL environment for every point")
    args = ap.parse_args()
    extra = dict(kv.split("=") for kv in args.env.split(",") if kv)
    for pt in args.points.split(","):
        f = pt.split(":")
        P, lanes, ctas, variant = int(f[0]), int(f[1]), int(f[2]), int(f[3])
        E = int(f[4]) if len(f) > 4 else args.E
        print(json.dumps(run_point(P, lanes, ctas, variant, E, args.regime, args.reps, extra)), flush=True)


if __name__ == "__main__":
    main()
