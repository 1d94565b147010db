#!/usr/bin/env python
"""Writes tests/golden/bench_state_gen30.npz: the openai_es state (mu, Adam m / v, sigma, t) that the
bench.py workload (BASELINE configs[2]: CartPole-v1 MLP, P = 65536, E = 5, sigma0 0.2, lr 0.1, seed 0)
reaches after 30 generations, computed on the CPU with the C bit-twin (oracle/ses_twin.c -- the GPU engine
reproduces its generations bit for bit, tests/test_gpu_loop.py).  From generation ~15 on every episode of
this run lasts the full 500 steps (the `history` array holds best reward and mean episode length per
generation); bench.py's timed generations 5..104 average 477 steps per episode.  bench.py's CPU legs start
the reference path from this state so that the CPU sample is timed in the regime the GPU arm is timed in,
instead of on generation-0 policies (episodes of ~20 steps, dominated by per-offspring overheads).

    python tools/make_bench_fixture.py [generations=30] [threads=os.cpu_count()]      (~4 min on 8 cores)
"""
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import twin  # noqa: E402

P, E, D, SEED, SIGMA0, DECAY, LR = 65536, 5, 226, 0, 0.2, 0.9999, 0.1


def adam_a(lr, t, beta1=0.99, beta2=0.999):
    """optimizers.py:44: stepsize * sqrt(1 - beta2^t) / (1 - beta1^t), as numpy evaluates it."""
    return lr * np.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)


def main():
    gens = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    nthreads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
    twin.build()
    mu = np.zeros(D, np.float32); m = np.zeros(D, np.float32); v = np.zeros(D, np.float32)
    sigma, t, hist = SIGMA0, 0, []
    for g in range(gens):
        t0 = time.time()
        fit, steps = twin.population_cartpole(mu[None], sigma=sigma, seed=SEED, gen=g, group=P, n_head=1, n=P, E=E,
                                              nthreads=nthreads)
        order = twin.rank_desc(fit)
        shaped = twin.centered_rank(order)
        grad = twin.grad_openai(shaped, D, SEED, g, P, 1, -(LR / (P * sigma)))
        t += 1
        mu, m, v = twin.adam(mu, m, v, grad, adam_a(LR, t))
        sigma *= DECAY
        hist.append((g, float(fit.max()), float(steps.sum()) / (P * E)))
        print("gen %d best %.1f mean episode length %.1f (%.1f s)" % (hist[-1] + (time.time() - t0,)), flush=True)
    out = os.path.join(ROOT, "tests", "golden", "bench_state_gen%d.npz" % gens)
    np.savez(out, mu=mu, m=m, v=v, sigma=np.float64(sigma), t=np.int64(t), generation=np.int64(gens),
             history=np.array(hist, dtype=np.float64))
    print("wrote", out)


if __name__ == "__main__":
    main()
