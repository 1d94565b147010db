#!/usr/bin/env python
"""Quickest possible A/B of K1 variants 4 and 6 in the converged regime (generation-30 state of the bench workload,
P = 65536, E = 5): 2 warm-up + 5 timed rollouts each, CUDA events, best of 5; checks that both give identical results."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simple_es_b200.engine import RolloutEngine  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "bench_state_gen30.npz"))
mu = torch.from_numpy(z["mu"][None].copy()).cuda()
sigma, P = float(z["sigma"]), 65536
out = {}
ref = None
for variant in (4, 6, 4, 6):
    os.environ["SES_K1_VARIANT"] = str(variant)
    eng = RolloutEngine("CartPole-v1", 4, 2, False, False, 500, 5, P, P, 1, 1, seed=0)
    best = 1e9
    for it in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fit, steps = eng.rollout(30, sigma, mu); e1.record(); torch.cuda.synchronize()
        if it >= 2:
            best = min(best, e0.elapsed_time(e1))
    if ref is None:
        ref = steps.clone()
    out.setdefault(variant, []).append(best)
    out["identical"] = bool(torch.equal(ref, steps)) and out.get("identical", True)
    eng.close()
out["env_steps"] = int(ref.sum())
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "k1_v6_quick.json"), "w").write(json.dumps(out) + "\n")
