#!/usr/bin/env python
"""Differential fuzzing of the kernel sources (host SIMT emulator, tests/simt_emu) against the C twin: random environment,
policy, population layout (group / heads / parents), slice, episode count, truncation, antithetic switch, init mode, grid
limits -- fitness and env-step counts must be bit-identical.  No GPU needed.

    python tools/emu_fuzz.py <first_case> <n_cases>          (16 260 cases incl. 3 000 with SES_K1_VARIANT=7 SES_GRU_VARIANT=1, 0 mismatches at the end of round 1)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from simt_emu.emu_engine import EmuEngine
from oracle import twin
twin.build()
seed0 = int(sys.argv[1]); n = int(sys.argv[2])
CLASSIC = {"MountainCar-v0": (2, 3), "Acrobot-v1": (6, 3)}
bad = 0
t0=time.time()
for case in range(seed0, seed0+n):
    rng = np.random.default_rng(case)
    kind = rng.choice(["mlp", "gru", "spread2", "spread3", "mc", "acro"])
    E = int(rng.choice([1, 2, 3, 4, 5, 6, 7, 9, 16, 32]))
    if kind == "gru": E = int(rng.choice([1,2,3,4,5,6,7,11]))
    P = int(rng.integers(2, 90 if kind in ("gru","acro") else 200))
    group = int(rng.choice([P, max(1, P // 3), 7])); group = max(1, min(group, P))
    n_par = (P - 1) // group + 1
    n_head = int(rng.integers(0, min(3, group) + 1))
    anti = bool(rng.random() < 0.3)
    init_mode = "fresh" if rng.random() < 0.5 else "shared"
    lo = int(rng.integers(0, P)); hi = int(rng.integers(lo, P + 1))
    if rng.random() < 0.5: lo, hi = 0, P
    os.environ["SES_ROLLOUT_CTAS_PER_SM"] = str(int(rng.integers(1, 3)))
    os.environ["SES_SIMT_EMU_SMS"] = str(int(rng.integers(1, 4)))
    if rng.random() < 0.3: os.environ["SES_ROLLOUT_LANES"] = str(int(rng.integers(1, 33)))
    else: os.environ.pop("SES_ROLLOUT_LANES", None)
    gen = int(rng.integers(0, 1000)); sigma = float(rng.choice([0.0, 0.1, 0.7, 2.0])); seed = int(rng.integers(0, 2**31))
    pomdp = bool(rng.random() < 0.3) and kind in ("mlp", "gru")
    twin.set_antithetic(anti)
    try:
        if kind in ("mlp", "gru"):
            ms = int(rng.choice([500, 40, 3]))
            eng = EmuEngine(population=P, group=group, n_head=n_head, n_parents=n_par, eval_ep_num=E, seed=seed, init_mode=init_mode,
                            id_begin=lo, id_end=hi, gru=(kind == "gru"), pomdp=pomdp, max_step=ms, antithetic=anti)
            par = rng.normal(0, 0.3, (n_par, eng.D)).astype(np.float32)
            fit, steps = eng.rollout(gen, sigma, par)
            tf, ts = twin.population_cartpole(par, gru=(kind == "gru"), pomdp=pomdp, sigma=sigma, seed=seed, gen=gen, group=group, n_head=n_head,
                                              id0=lo, n=hi - lo, E=E, max_step=ms, init_mode=0 if init_mode == "shared" else 1, nthreads=2)
        elif kind.startswith("spread"):
            N = int(kind[-1])
            eng = EmuEngine(env_name="simple_spread", obs_dim=6 * N, act_dim=5, n_agents=N, max_step="None", population=P, group=group, n_head=n_head,
                            n_parents=n_par, eval_ep_num=E, seed=seed, init_mode=init_mode, id_begin=lo, id_end=hi, antithetic=anti)
            par = rng.normal(0, 0.5, (n_par, eng.D)).astype(np.float32)
            fit, steps = eng.rollout(gen, sigma, par)
            tf, ts = twin.population_mpe(par, N=N, sigma=sigma, seed=seed, gen=gen, group=group, n_head=n_head, id0=lo, n=hi - lo, E=E,
                                         init_mode=0 if init_mode == "shared" else 1)
        else:
            env = "MountainCar-v0" if kind == "mc" else "Acrobot-v1"
            obs, act = CLASSIC[env]
            ms = int(rng.choice([60, 25, 3])) if kind == "acro" else int(rng.choice([200, 50]))
            eng = EmuEngine(env_name=env, obs_dim=obs, act_dim=act, max_step=ms, population=P, group=group, n_head=n_head, n_parents=n_par,
                            eval_ep_num=E, seed=seed, init_mode=init_mode, id_begin=lo, id_end=hi, antithetic=anti)
            par = rng.normal(0, 0.5, (n_par, eng.D)).astype(np.float32)
            fit, steps = eng.rollout(gen, sigma, par)
            tf, ts = twin.population_classic(env, par, sigma=sigma, seed=seed, gen=gen, group=group, n_head=n_head, id0=lo, n=hi - lo, E=E,
                                             max_step=ms, init_mode=0 if init_mode == "shared" else 1, nthreads=2)
        ok = np.array_equal(steps[lo:hi], ts) and np.array_equal(fit[lo:hi], tf) and np.all(steps[:lo] == -1) and np.all(steps[hi:] == -1)
        eng.close()
    except Exception as exc:
        ok = False; print("EXC", case, kind, repr(exc))
    finally:
        twin.set_antithetic(False)
    if not ok:
        bad += 1
        print("MISMATCH case", case, kind, dict(E=E, P=P, group=group, n_head=n_head, anti=anti, init=init_mode, lo=lo, hi=hi, gen=gen, sigma=sigma, pomdp=pomdp))
print("cases", n, "bad", bad, "time %.1f" % (time.time() - t0))
