#!/usr/bin/env python
"""Scaling sweep of any config under torchrun (BASELINE configs[4]: CartPole simple_genetic, P = 2^16 .. 2^20 on 1..8 GPUs;
also weak scaling of the headline config).  Rank 0 prints one JSON line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 \\
        tools/scale_bench.py --conf cartpole_genetic.yaml --offspring-num 1048576 --elite-num 16 --generations 20
"""
import argparse
import json
import os
import sys

import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simple_es_b200 import dist as sdist  # noqa: E402
from simple_es_b200.loop import B200Loop  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--conf", default="cartpole_genetic.yaml")
    ap.add_argument("--offspring-num", type=int, default=1 << 20)
    ap.add_argument("--elite-num", type=int, default=None)
    ap.add_argument("--generations", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--shard", default="cyclic", choices=["cyclic", "contiguous"])
    args = ap.parse_args()
    rank, world = sdist.init_from_env()
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    cfg = yaml.load(open(os.path.join(ROOT, "conf", args.conf)), Loader=yaml.FullLoader)
    cfg["strategy"]["offspring_num"] = args.offspring_num
    if args.elite_num is not None:
        cfg["strategy"]["elite_num"] = args.elite_num
    cfg["engine"]["shard"] = args.shard
    loop = B200Loop(cfg, args.generations, 1, 5, save_model_period=0, seed=0, device=local, quiet=True)
    s = loop.strategy
    for _ in range(args.warmup):
        s.step()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    n0 = int(s.total_env_steps.item())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.generations):
        s.step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    n = torch.tensor([int(s.total_env_steps.item()) - n0], dtype=torch.int64, device="cuda")
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(n, op=torch.distributed.ReduceOp.SUM)
    best = float(s.best_reward().item())
    if world > 1:
        if s.exchange == "peer":
            s.engine.peer_check()
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank == 0:
        ms = float(t[0])
        print(json.dumps({"conf": args.conf, "strategy": cfg["strategy"]["name"], "population": s.P, "n_gpus": world, "shard": args.shard,
                          "generations": args.generations, "ms_per_generation": ms / args.generations,
                          "generations_per_s": args.generations / (ms * 1e-3), "env_steps_per_s": int(n[0]) / (ms * 1e-3),
                          "best_reward": best}))


if __name__ == "__main__":
    main()
