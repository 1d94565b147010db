#!/usr/bin/env python
"""Dynamic operation counts per env step of every rollout kernel, taken from the kernel SOURCE on the host SIMT emulator
(tests/simt_emu, counting build -DSIMT_EMU_COUNT): the number of arithmetic intrinsics one lane executes per env step.
No GPU, no profiler.  Method: the same population is rolled out twice with different step limits (every episode runs to the
limit), and the counter difference is divided by the difference in env steps -- offspring set-up, weight regeneration and
scheduling cancel exactly.  Packed FFMA2 / FMUL2 / FADD2 count as ONE packed instruction (= two lane operations).

    python tools/emu_opcount.py            -> profiles/r01_emu_opcounts.md
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from simt_emu import build as b  # noqa: E402

NAMES = ["f32_fma", "f32_mul", "f32_add", "f32x2_fma", "f32x2_mul", "f32x2_add", "f32_div", "f32_sqrt", "f32_rcp", "f32_minmax",
         "f64_fma", "f64_mul", "f64_add", "f64_div", "f64_sqrt", "f64_rcp", "shfl", "vote", "match", "syncwarp", "syncthreads", "atomic"]


def build_counting():
    b.build()
    lib = os.path.join(b.BUILD, "libses_simt_emu_count.so")
    subprocess.check_call(["g++"] + b.CXXFLAGS + ["-DSIMT_EMU_COUNT", "-I", os.path.join(b.HERE, "include"), "-o", lib,
                                                  os.path.join(b.BUILD, "src", "ses_abi.cpp")])
    return lib


def main():
    os.environ["SES_SIMT_EMU_LIB"] = build_counting()
    from simt_emu import emu_engine
    lib = emu_engine.load()
    lib.simt_emu_counters.argtypes = [C.c_void_p, C.c_int]

    def counters(reset=True):
        out = np.zeros(len(NAMES), dtype=np.uint64)
        lib.simt_emu_counters(out.ctypes.data, int(reset))
        return out.astype(np.int64)

    def balancing(D, gru):
        mu = np.zeros((1, D), np.float32)
        if gru:
            W1 = mu[0, 0:128].reshape(32, 4); Wih = mu[0, 160:160 + 3072].reshape(96, 32); bih = mu[0, 160 + 6144:160 + 6144 + 96]
            W2 = mu[0, 160 + 6144 + 192:160 + 6144 + 192 + 64].reshape(2, 32)
            W1[0] = [0.0, 0.5, 10.0, 3.0]; Wih[64, 0] = 3.0; bih[32:64] = -10.0; W2[1, 0] = 5.0; W2[0, 0] = -5.0
        else:
            w1 = mu[0, :128].reshape(32, 4); w2 = mu[0, 160:224].reshape(2, 32)
            w1[0] = [0.0, 0.5, 10.0, 3.0]; w2[1, 0] = 5.0; w2[0, 0] = -5.0
        return mu

    def per_step(make, mu_fn, sigma, limits, note):
        res = []
        for ms in limits:
            eng = make(ms)
            mu = mu_fn(eng.D)
            counters()
            fit, steps = eng.rollout(3, sigma, mu)
            res.append((counters(), int(steps.sum())))
            eng.close()
        (c0, n0), (c1, n1) = res
        assert n1 > n0
        return (c1 - c0) / float(n1 - n0), n1 - n0, note

    rows = []
    for v in (0, 2, 4, 7, 6):
        os.environ["SES_K1_VARIANT"] = str(v)
        rows.append(("CartPole-v1 MLP, K1 variant %d" % v,) + per_step(
            lambda ms: emu_engine.EmuEngine(population=30, group=30, n_head=1, eval_ep_num=5, seed=1, max_step=ms),
            lambda D: balancing(D, False), 0.02, (100, 200), "env step = 1 episode step of one lane"))
    os.environ["SES_K1_VARIANT"] = "4"
    for v in (0, 1):
        os.environ["SES_GRU_VARIANT"] = str(v)
        rows.append(("CartPole-v1 GRU (5 episodes per warp), variant %d" % v,) + per_step(
            lambda ms: emu_engine.EmuEngine(population=8, group=8, n_head=2, eval_ep_num=5, seed=1, max_step=ms, gru=True),
            lambda D: balancing(D, True), 0.01, (60, 120), "per env step, summed over the warp's 32 lanes (lane = hidden unit)"))
    os.environ["SES_GRU_VARIANT"] = "0"
    for N in (2, 3):
        Dn = 6 * N * 32 + 32 + 165
        rows.append(("simple_spread N = %d" % N,) + per_step(
            lambda ms: emu_engine.EmuEngine(env_name="simple_spread", obs_dim=6 * N, act_dim=5, n_agents=N, max_step=ms, population=30,
                                            group=30, n_head=1, eval_ep_num=5, seed=1),
            lambda D: np.random.default_rng(0).normal(0, 0.5, (1, D)).astype(np.float32), 0.3, (10, 20), "world step = N agent actions"))
    rows.append(("MountainCar-v0",) + per_step(
        lambda ms: emu_engine.EmuEngine(env_name="MountainCar-v0", obs_dim=2, act_dim=3, max_step=ms, population=30, group=30, n_head=1,
                                        eval_ep_num=5, seed=1),
        lambda D: np.zeros((1, D), np.float32), 0.0, (50, 100), "all-zero policy: never reaches the goal"))
    out = ["# Dynamic operation counts per env step, from the kernel sources on the host SIMT emulator (`tools/emu_opcount.py`)", "",
           "Arithmetic intrinsics executed per env step (difference of two runs with different step limits: set-up and weight",
           "regeneration cancel).  `f32x2_*` are packed instructions (two float32 lane operations each).  FP32 FMA-pipe lane operations",
           "= f32_fma + f32_mul + f32_add + 2 x (f32x2_*); float64 = f64_fma + f64_mul + f64_add (+ the expansion of f64_div / f64_sqrt,",
           "~10 more each on the device).", "",
           "| kernel | " + " | ".join(NAMES[:16]) + " | shfl | FP32 lane-ops | FP64 ops | note |", "|---|" + "---|" * 20]
    for name, c, nsteps, note in rows:
        d = dict(zip(NAMES, c))
        f32 = d["f32_fma"] + d["f32_mul"] + d["f32_add"] + 2 * (d["f32x2_fma"] + d["f32x2_mul"] + d["f32x2_add"])
        f64 = d["f64_fma"] + d["f64_mul"] + d["f64_add"]
        out.append("| %s | " % name + " | ".join("%.1f" % d[k] if d[k] % 1 else "%d" % d[k] for k in NAMES[:16]) +
                   " | %.1f | %.0f | %.0f | %s (%d steps) |" % (d["shfl"], f32, f64, note, nsteps))
    text = "\n".join(out) + "\n"
    open(os.path.join(ROOT, "profiles", "r01_emu_opcounts.md"), "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
