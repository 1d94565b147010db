#!/usr/bin/env python
"""Whole-generation time (K1 + exchange + K2 + K3, strategy.step()) of every BASELINE config on one GPU, with the K1
share, from CUDA events.  Prints one JSON line per config."""
import json
import os
import sys

import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simple_es_b200.loop import B200Loop  # noqa: E402

CASES = [  # label, conf, strategy overrides, generations (warm-up 3)
    ("C0 cartpole.yaml as shipped (P=97)", "cartpole.yaml", {}, 40),
    ("C1 CartPole POMDP GRU simple_evolution P=4097", "cartpole_pomdp_gru.yaml", {"offspring_num": 4096}, 30),
    ("C2 CartPole MLP openai_es P=65536", "cartpole_openai.yaml", {"offspring_num": 65536}, 30),
    ("C3 simple_spread N=3 openai_es P=16384", "simplespread.yaml", {"offspring_num": 16384}, 30),
    ("C4 CartPole simple_genetic P=2^20", "cartpole_genetic.yaml", {"offspring_num": 1 << 20, "elite_num": 16}, 10),
    ("MountainCar-v0 simple_genetic P=16384", "mountaincar.yaml", {}, 20),
    ("Acrobot-v1 openai_es P=16384", "acrobot.yaml", {}, 20),
]


def main():
    for label, conf, over, gens in CASES:
        cfg = yaml.load(open(os.path.join(ROOT, "conf", conf)), Loader=yaml.FullLoader)
        cfg["strategy"].update(over)
        if conf == "simplespread.yaml":
            cfg["network"]["num_state"] = 18; cfg["engine"]["n_agents"] = 3
        loop = B200Loop(cfg, gens, 1, 5, save_model_period=0, seed=0, quiet=True)
        s = loop.strategy
        for _ in range(3):
            s.step()
        torch.cuda.synchronize()
        steps0 = int(s.total_env_steps.item())
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(gens + 1)]
        ev[0].record()
        for g in range(gens):
            s.step()
            ev[g + 1].record()
        torch.cuda.synchronize()
        ms = [ev[g].elapsed_time(ev[g + 1]) for g in range(gens)]
        n = int(s.total_env_steps.item()) - steps0
        # K1 alone on the final parents
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.engine.rollout(s.generation, s.sigma, s.parents, fitness=s.fitness, steps=s.steps); e1.record()
        torch.cuda.synchronize()
        k1 = e0.elapsed_time(e1)
        tot = sum(ms)
        print(json.dumps({"config": label, "ms_per_generation": tot / gens, "last_generation_ms": ms[-1], "k1_ms_last": k1,
                          "generations_per_s": gens / (tot * 1e-3), "env_steps_per_s": n / (tot * 1e-3),
                          "best_reward": float(s.best_reward().item())}))
        del loop, s
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
