#!/usr/bin/env python
"""Calibration of bench.py's reference arm: the UNMODIFIED reference `ESLoop` (imported from /root/reference
through oracle/ref_bridge.py -- dev container only) timed next to the port bench.py runs (oracle/pyref.py::es_loop_port) on the
same population sample, the same cores, the same regime (the committed generation-30 state of the bench workload).

The reference's own loop deep-copies and pickles one nn.Module per offspring and builds the next population as modules; the port
ships flat weight vectors.  The ratio written here says by how much the port flatters the reference.

    python tools/calibrate_reference.py [offspring=365] [generations=2]     ->  profiles/r02_reference_calibration.json
"""
import contextlib
import io
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", "1")
import bench  # noqa: E402
from oracle import pyref, ref_bridge  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 365
    gens = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    cores = os.cpu_count() or 1
    torch.set_num_threads(1)
    ref = ref_bridge.load()
    st = bench.bench_state()
    E = bench.E_DEFAULT
    init = np.random.RandomState(0).uniform(-0.05, 0.05, size=(E, 4))

    # ---- the reference's own ESLoop + openai_es + GymEnvModel + RolloutWorker over the gym-free CartPole shim
    env = pyref.CartPoleShim(max_step=500, init_states=init)
    env.name = "CartPole-v1"
    net = ref.GymEnvModel(4, 2, True, False)
    strat = ref.openai_es(st["sigma"], bench.STRATEGY["sigma_decay"], bench.STRATEGY["learning_rate"], n)
    steps_ref = []
    orig_worker = ref.RolloutWorker
    with tempfile.TemporaryDirectory() as tmp:
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            loop = ref.ESLoop({}, strat, env, net, gens, cores, E, log=False, save_model_period=10 ** 9)
            # Start the UNMODIFIED classes from the trained state.  ESLoop.__init__ (loop.py:31) and openai_es.init_offspring
            # (offspring_strategies.py:348) both zero the network, so after init_offspring has run, mu is loaded into the
            # strategy's mu_model, the first population is drawn again by the reference's own _gen_offsprings, and the Adam
            # object init_offspring created receives the state's moments and step count before the first evaluate().
            from simple_es_b200 import checkpoint
            mu_sd = checkpoint.flat_to_state_dict(torch.from_numpy(st["mu"]), 4, 2, False)
            shapes = pyref.param_shapes(4, 2, False)
            orig_init, orig_eval = strat.init_offspring, strat.evaluate

            def init_with_state(network, agent_ids):
                orig_init(network, agent_ids)
                strat.mu_model.load_state_dict(mu_sd)
                pop = strat._gen_offsprings(strat.agent_ids, strat.mu_model, strat.curr_sigma, strat.offspring_num)
                strat.optimizer.m = [a.copy() for a in pyref.flat_to_list(st["m"], 4, 2, False)]
                strat.optimizer.v = [a.copy() for a in pyref.flat_to_list(st["v"], 4, 2, False)]
                strat.optimizer.t = int(st["t"])
                return pop

            def eval_counting(rewards):
                steps_ref.append(float(np.sum(rewards)) * E)          # CartPole: reward 1 per step, fitness = steps / E
                return orig_eval(rewards)
            strat.init_offspring, strat.evaluate = init_with_state, eval_counting
            np.random.seed(12345)
            out = io.StringIO()
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(out):
                loop.run()
            t_ref = time.perf_counter() - t0
        finally:
            os.chdir(cwd)
    lines = [l for l in out.getvalue().splitlines() if l.startswith("episode")]
    best = [float(l.split("Best reward:")[1].split(",")[0]) for l in lines]
    # ---- the port, same sample
    t0 = time.perf_counter()
    tot = 0
    for g in range(gens):
        s, dt = bench.cpu_reference_generation(n, cores, 12345 + g, state=st)
        tot += s
    t_port = time.perf_counter() - t0
    ref_steps = int(round(sum(steps_ref)))
    rec = {"offspring": n, "generations": gens, "cores": cores, "eval_ep_num": E,
           "reference_ESLoop": {"seconds": t_ref, "env_steps": ref_steps, "env_steps_per_s": ref_steps / t_ref, "best_rewards": best,
                                "what": "unmodified learning_strategies/evolution/loop.py ESLoop.run + openai_es + GymEnvModel + RolloutWorker, "
                                        "mp.Pool(%d) per generation, oracle/pyref.py::CartPoleShim behind the GymWrapper duck type" % cores},
           "port_es_loop_port": {"seconds": t_port, "env_steps": tot, "env_steps_per_s": tot / t_port,
                                 "what": "oracle/pyref.py::es_loop_port as bench.py --impl reference / cpu_baseline run it"},
           "port_over_reference": (tot / t_port) / (ref_steps / t_ref),
           "where": "dev container (no GPU), %d cores" % cores}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    path = os.path.join(ROOT, "profiles", "r02_reference_calibration.json")
    with open(path, "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
