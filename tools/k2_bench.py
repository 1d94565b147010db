#!/usr/bin/env python
"""K2 A/B on one GPU: ses_rank_desc (rank + centered ranks) with the fused build (SES_K2_PERSISTENT=0: 1 + passes launches) and
the persistent kernel (one cooperative launch, grid barriers between the passes).  CUDA events, 5 warm-up + 50 timed calls,
microseconds per call.

    python tools/k2_bench.py            (writes gpurun_out/k2_ab.jsonl)
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simple_es_b200.engine import RolloutEngine  # noqa: E402

CASES = [("CartPole-v1 integer keys (2 passes)", 65536, "int"), ("CartPole-v1 integer keys, one rank's view at N=8 (same vector)", 65536, "int_conv"),
         ("simple_spread float64 keys (8 passes)", 16384, "float"), ("genetic 2^20 integer keys", 1 << 20, "int")]


def main():
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = open(os.path.join(ROOT, "gpurun_out", "k2_ab.jsonl"), "w")
    rng = np.random.default_rng(0)
    for name, P, kind in CASES:
        if kind == "float":
            fit = torch.from_numpy(-rng.uniform(10, 60, P)).cuda()
        elif kind == "int_conv":
            fit = torch.full((P,), 500.0, dtype=torch.float64, device="cuda")          # converged population: every key equal
        else:
            fit = torch.from_numpy(rng.integers(40, 2501, P) / 5.0).cuda()
        res = {}
        for fused in (0, 1):
            os.environ["SES_K2_PERSISTENT"] = str(fused)
            eng = RolloutEngine("CartPole-v1", 4, 2, False, False, 500, 5, P, P, 1, 1, seed=0)
            order = torch.empty(P, dtype=torch.int32, device="cuda"); shaped = torch.empty(P, dtype=torch.float64, device="cuda")
            full = kind == "float"
            for _ in range(5):
                eng.rank_desc(fit, shaped=True, order=order, shaped_out=shaped, full_key=full)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = eng.launches
            e0.record()
            for _ in range(50):
                eng.rank_desc(fit, shaped=True, order=order, shaped_out=shaped, full_key=full)
            e1.record(); torch.cuda.synchronize()
            res[fused] = (e0.elapsed_time(e1) * 1e3 / 50, (eng.launches - l0) // 50, order.clone(), shaped.clone())
            eng.close()
        same = bool(torch.equal(res[0][2], res[1][2]) and torch.equal(res[0][3], res[1][3]))
        line = {"case": name, "P": P, "fused_us": res[0][0], "fused_launches": res[0][1], "persistent_us": res[1][0],
                "persistent_launches": res[1][1], "identical_outputs": same}
        print(json.dumps(line)); out.write(json.dumps(line) + "\n")
    out.close()


if __name__ == "__main__":
    main()
