#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 population-rollout engine.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json metric: "env-steps/sec/GPU, CartPole-v1 pop=65536 ...; ES generations/sec"):
CartPole-v1, 32-hidden MLP policy (D = 226), openai_es (centered ranks + Adam), population 65536
sharded over N GPUs (strong scaling), eval_ep_num 5, max_step 500, seed 0, mu = 0 at generation 0
(conf/cartpole_openai.yaml).  A "step" is one ES generation: K1 rollout of the rank's slice ->
fitness all-gather (N > 1) -> K2 rank/shape -> K3 gradient + Adam.  `value` counts env steps that were
actually simulated (sum of episode lengths), not P*E*500.

Prints ONE JSON line on rank 0 (see DESIGN.md section 8 for every key).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# One OpenMP thread per process (what torchrun sets for N > 1 anyway): the CPU legs run one forked worker per
# core, and torch's intra-op pool inside each of them only oversubscribes the box (measured here: 116 k -> 180 k
# env-steps/s for the reference path on 8 cores).  The GPU arm does no CPU compute.
os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("MKL_NUM_THREADS", "1")

P_DEFAULT = 65536
E_DEFAULT = 5
D = 226
FLOP_PER_STEP = 418          # 192 FMA + 34 bias adds per env step (CartPole MLP)
FMA_LANE_OPS_PER_STEP = 640  # executed on the FP32 FMA pipe per env step: fc1 128 + fc2 64 + 32 tanh x 14
K1_DRAM_BYTES_PER_LAUNCH = 33280   # ncu, profiles/r01_k1_v4_conv.txt (reads; no DRAM writes: results stay in L2)
STRATEGY = dict(name="openai_es", init_sigma=0.2, sigma_decay=0.9999, learning_rate=0.1)


# ----------------------------------------------------------------------------------------------
# clocks (B200_PROFILING.md "clocks DURING the timed region")
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock + throttle reasons every ~5 ms through NVML (nvidia_ml_py) while the timed
    region runs; the same fields as the recipe's nvidia-smi clocks line."""
    BITS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
            "hw_power_brake_slowdown": 0x80}

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.sm, self.reasons, self.power = [], set(), []
        self.max_mhz = None
        self._stop = threading.Event()
        self.thread = None
        self.err = None

    def _phys_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.gpu])
            except Exception:
                return self.gpu
        return self.gpu

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._phys_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as exc:
            self.err = "nvml unavailable: %s" % (exc,)
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _sample(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for name, bit in self.BITS.items():
            if r & bit:
                self.reasons.add(name)
        try:
            self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception as exc:
                self.err = str(exc)
                return
            time.sleep(0.005)

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "no sampler"]}
        try:
            self._sample()
        except Exception:
            pass
        self._stop.set()
        self.thread.join(timeout=1)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(sm), "power_w_max": max(self.power) if self.power else None}


# ----------------------------------------------------------------------------------------------
# CPU legs (oracle; allowed importers of oracle/: tests, smoke, and these two functions)
# ----------------------------------------------------------------------------------------------
STATE_FIXTURE = os.path.join(ROOT, "tests", "golden", "bench_state_gen30.npz")
REGIME = ("population drawn around the generation-30 state of this very workload (tests/golden/bench_state_gen30.npz, "
          "tools/make_bench_fixture.py): episodes of ~500 steps, the regime the GPU arm's timed generations run in")
CPU_SEC_PER_OFFSPRING = 0.21   # 5 episodes x 500 steps x ~85 us per reference policy-forward + env step on one core


def bench_state():
    """The trained openai_es state the CPU legs start from (a committed fixture; not an oracle import)."""
    import numpy as np
    z = np.load(STATE_FIXTURE)
    return {"mu": z["mu"], "m": z["m"], "v": z["v"], "sigma": float(z["sigma"]), "t": int(z["t"])}


def cpu_reference_generation(n_offspring, cores, gen_seed, eval_ep_num=E_DEFAULT, state=None):
    """One generation of the reference's CPU path (port: oracle/pyref.py -- torch-CPU policy per
    offspring, Python CartPole, a fresh mp.Pool(cores), strategy.evaluate) on a population sample of
    n_offspring drawn around `state` (None: the all-zero generation-0 network).  Returns (env_steps, seconds)."""
    import numpy as np
    import torch
    from oracle import pyref
    torch.set_num_threads(1)
    np.random.seed(gen_seed)
    init = np.random.RandomState(0).uniform(-0.05, 0.05, size=(eval_ep_num, 4))
    env = pyref.CartPoleShim(max_step=500, init_states=init)
    cfg = dict(STRATEGY, offspring_num=n_offspring)
    t0 = time.perf_counter()
    rec = pyref.es_loop_port(env, (4, 2, False), cfg, 1, cores, eval_ep_num, seed=gen_seed, state=state)
    return rec[0]["env_steps"], time.perf_counter() - t0


def cpu_baseline_block(cores):
    """Bounded CPU baseline reported beside the GPU number (rank 0, N = 1)."""
    n = max(64, min(P_DEFAULT, 96 * cores))            # 10-20 s of CPU work on the box's cores
    steps, dt = cpu_reference_generation(n, cores, 12345, state=bench_state())
    out = {"value": steps / dt, "unit": "env-steps/s", "cores": cores, "kind": "port",
           "sample": "1 generation of the reference CPU path (oracle/pyref.py port: torch-CPU policy, Python CartPole, "
                     "mp.Pool(%d), evaluate) on %d offspring x %d episodes = %d env steps in %.1f s; %s"
                     % (cores, n, E_DEFAULT, steps, dt, REGIME)}
    try:  # a much stronger CPU number for context: the C bit-twin on every host thread, same regime
        import numpy as np
        from oracle import twin
        st = bench_state()
        n2 = 512 * cores
        t0 = time.perf_counter()
        _, ts = twin.population_cartpole(st["mu"][None], sigma=st["sigma"], seed=0, gen=30, group=P_DEFAULT, n_head=1,
                                         n=n2, E=E_DEFAULT, nthreads=cores)
        dt2 = time.perf_counter() - t0
        out["c_twin"] = {"value": float(ts.sum()) / dt2, "unit": "env-steps/s", "cores": cores,
                         "sample": "%d offspring of generation 30 x %d episodes = %d env steps, oracle/ses_twin.c, %d pthreads"
                                   % (n2, E_DEFAULT, int(ts.sum()), cores)}
    except Exception as exc:  # pragma: no cover
        out["c_twin"] = {"error": str(exc)}
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path, timed on the host cores."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    # per-step sample: the whole run (K timed + W warm-up steps) should take about two minutes on this box
    per_step_s = 120.0 / max(1, args.steps + args.warmup)
    n = int(max(2 * cores, min(64 * cores, per_step_s * cores / CPU_SEC_PER_OFFSPRING)))
    state = bench_state()
    for w in range(args.warmup):
        cpu_reference_generation(n, cores, 100 + w, state=state)
    tot_steps, tot_t = 0, 0.0
    for k in range(args.steps):
        s, dt = cpu_reference_generation(n, cores, 1000 + k, state=state)
        tot_steps += s; tot_t += dt
    v = tot_steps / tot_t
    sample = ("each step = 1 generation of the reference CPU path (oracle/pyref.py port, mp.Pool(%d)) on %d offspring x %d "
              "episodes (bounded sample of the 65536 population); %s" % (cores, n, E_DEFAULT, REGIME))
    line = {
        "impl": "reference", "metric": "env-steps/sec", "value": v, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, args.steps),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 policy / f64 physics",
        "data": "synthetic", "config": workload_config(args, 1),
        "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(args, world):
    return {"workload": "CartPole-v1 MLP(4-32-2, D=226) openai_es pop=%d eval_ep_num=%d max_step=500 "
                        "(BASELINE configs[2]; conf/cartpole_openai.yaml)" % (args.pop, E_DEFAULT),
            "population": args.pop, "eval_ep_num": E_DEFAULT, "strategy": STRATEGY,
            "parallelism": "pop-shard x%d (%s)" % (world, args.shard if world > 1 else "single"), "fitness_exchange": (args.exchange if world > 1 else "local"),
            "init_states": "shared [E] table (reference mp.Pool semantics)",
            "l2": "flushed between timed generations (256 MiB write); the hot path itself reads 904 B of parameters per generation"}


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from simple_es_b200 import dist as sdist
    from simple_es_b200.loop import B200Loop

    rank, world = sdist.init_from_env()
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    config = {"env": {"name": "CartPole-v1", "max_step": 500, "pomdp": False},
              "network": {"name": "gym_model", "num_state": 4, "num_action": 2, "discrete_action": True, "gru": False},
              "strategy": dict(STRATEGY, offspring_num=args.pop),
              "engine": {"name": "b200", "fitness_exchange": args.exchange, "shard": args.shard}}
    loop = B200Loop(config, args.steps + args.warmup, 1, E_DEFAULT, log=False, save_model_period=0, seed=0, device=local, quiet=True)
    s = loop.strategy
    eng = s.engine
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        s.step()
    barrier()
    steps_before = int(s.total_env_steps.item())
    launches_before = eng.launches
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)                      # L2 flush, outside the timed events
        ev[k][0].record()
        # one generation, with an extra event after K1 so that the dominant kernel is timed alone
        e = s.engine
        s.fitness = s._fit[s.generation & 1]
        e.rollout(s.generation, s.sigma, s.parents, fitness=s.fitness, steps=s.steps)
        ev[k][1].record()
        s.exchange_fitness()
        e.rank_desc(s.fitness, shaped=True, order=s.order, shaped_out=s.shaped)
        s.t += 1
        e.update_openai(s.generation, s.sigma, s.lr, s.t, s.shaped, s.parents.view(-1), s.m, s.v)
        s.sigma *= s.decay; s.curr_sigma = s.sigma; s.generation += 1
        ev[k][2].record()
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    gen_ms = sum(ev[k][0].elapsed_time(ev[k][2]) for k in range(args.steps))
    k1_ms = sum(ev[k][0].elapsed_time(ev[k][1]) for k in range(args.steps))
    local_steps = int(s.total_env_steps.item()) - steps_before
    k1_local_steps = local_steps
    best = float(s.best_reward().item())
    launches = eng.launches - launches_before
    t = torch.tensor([gen_ms, k1_ms], dtype=torch.float64, device=dev)
    tsum = t.clone()
    n = torch.tensor([local_steps], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
    # load balance over ranks: K1 time of the slowest rank vs the mean
    k1_rank_ms = {"max": float(t[1]) / args.steps, "mean": float(tsum[1]) / world / args.steps}
    gen_ms, k1_ms = float(t[0]), float(t[1])
    total_steps = int(n[0])

    # ---------------- e2e: the same generations with HOST buffers, copies inside the timed region (wall clock, max over ranks)
    e2e = None
    if world == 1:
        # the reference-facing C-ABI call: ses_generation_openai_host (H2D mu/m/v, K1-K3, D2H fitness[P] + mu/m/v + steps)
        import numpy as np
        from simple_es_b200.engine import RolloutEngine
        P = args.pop
        eng2 = RolloutEngine("CartPole-v1", 4, 2, False, False, 500, E_DEFAULT, P, P, 1, 1, seed=0, device=local)
        pin = lambda n_, dt: torch.empty(n_, dtype=dt).pin_memory().numpy()
        mu, m, v = pin(D, torch.float32), pin(D, torch.float32), pin(D, torch.float32)
        fit = pin(P, torch.float64)
        mu[:] = 0; m[:] = 0; v[:] = 0
        sigma = STRATEGY["init_sigma"]
        for g in range(args.warmup):
            eng2.generation_openai_host(g, sigma, STRATEGY["learning_rate"], g + 1, mu, m, v, fit)
            sigma *= STRATEGY["sigma_decay"]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tot = 0
        for g in range(args.warmup, args.warmup + args.steps):
            tot += eng2.generation_openai_host(g, sigma, STRATEGY["learning_rate"], g + 1, mu, m, v, fit)
            sigma *= STRATEGY["sigma_decay"]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e = {"value": tot / dt, "unit": "env-steps/s", "h2d_bytes_per_step": 3 * D * 4, "d2h_bytes_per_step": P * 8 + 3 * D * 4 + 8,
               "api": "ses_generation_openai_host (C ABI, pinned host buffers, synchronous)", "n_gpus": 1, "env_steps": tot,
               "generations_per_sec": args.steps / dt}
        eng2.close()
    else:
        # N ranks: every rank stages mu/m/v from pinned host memory, runs its shard of the generation through the
        # public strategy API (B200Loop.strategy.step()), and reads the full fitness vector + mu/m/v back to the host
        P = args.pop
        pinned = lambda n_, dt: torch.empty(n_, dtype=dt).pin_memory()
        mu_h, m_h, v_h, fit_h = pinned(D, torch.float32), pinned(D, torch.float32), pinned(D, torch.float32), pinned(P, torch.float64)
        mu_h.copy_(s.parents[0]); m_h.copy_(s.m); v_h.copy_(s.v)
        torch.cuda.synchronize()
        steps0 = int(s.total_env_steps.item())
        barrier()
        t0 = time.perf_counter()
        for k in range(args.steps):
            s.parents[0].copy_(mu_h, non_blocking=True); s.m.copy_(m_h, non_blocking=True); s.v.copy_(v_h, non_blocking=True)
            s.step()
            fit_h.copy_(s.fitness, non_blocking=True)
            mu_h.copy_(s.parents[0], non_blocking=True); m_h.copy_(s.m, non_blocking=True); v_h.copy_(s.v, non_blocking=True)
            torch.cuda.synchronize()
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        nn = torch.tensor([int(s.total_env_steps.item()) - steps0], dtype=torch.int64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(nn, op=dist.ReduceOp.SUM)
        e2e = {"value": int(nn[0]) / float(tt[0]), "unit": "env-steps/s", "h2d_bytes_per_step": 3 * D * 4,
               "d2h_bytes_per_step": P * 8 + 3 * D * 4,
               "api": "B200Loop.strategy.step() per rank with pinned host staging of mu/m/v (H2D) and fitness[P] + mu/m/v (D2H), "
                      "synchronous; bytes are per rank", "n_gpus": world, "env_steps": int(nn[0]),
               "generations_per_sec": args.steps / float(tt[0])}

    if world > 1:
        s.engine.peer_check() if s.exchange == "peer" else None
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    peak_tf = measure_fp32_peak(local)
    achieved_tf = k1_local_steps * FLOP_PER_STEP / (k1_ms * 1e-3) / 1e12 if k1_ms > 0 else 0.0
    value = total_steps / (gen_ms * 1e-3)
    line = {
        "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": gen_ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32 policy / f64 physics", "data": "synthetic",
        "config": workload_config(args, world),
        "per_gpu": value / world, "generations_per_sec": args.steps / (gen_ms * 1e-3), "env_steps": total_steps,
        "best_reward_last_gen": best, "wall_s": wall, "k1_ms_per_generation_over_ranks": k1_rank_ms,
        "roofline": {"bound": "fp32_pipe", "kernel": "k_rollout_cartpole_mlp", "achieved": achieved_tf, "peak": peak_tf,
                     "unit": "TFLOP/s", "frac": achieved_tf / peak_tf if peak_tf else None, "traffic": K1_DRAM_BYTES_PER_LAUNCH,
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of K1 (profiles/r01_k1_v4_conv.txt); "
                                       "algorithmic bytes per launch: 904 B of parameters + 16 B per offspring of results",
                     "peak_source": "measured live: dependent-free FFMA microbenchmark on this GPU (MEASURED_PEAKS.json has only HBM and bf16-tensor peaks)",
                     "algorithmic": "%d FP32 FLOP per env step (SURVEY 8d) x %d env steps of rank 0 / %.3f ms in K1" % (FLOP_PER_STEP, k1_local_steps, k1_ms),
                     "k1_share_of_step": k1_ms / gen_ms if gen_ms else None,
                     "fma_pipe_frac": (k1_local_steps * FMA_LANE_OPS_PER_STEP * 2 / (k1_ms * 1e-3) / 1e12 / peak_tf) if (peak_tf and k1_ms > 0) else None,
                     "fma_pipe_note": "share of the FP32 FMA pipe K1 keeps busy: 640 FMA-pipe lane operations per env step (the 418 algorithmic "
                                      "FLOP count no tanh; an accurate float32 tanh costs 14 FMA-pipe operations) x 2 FLOP / measured FFMA peak"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
    }
    line["roofline"]["hbm_check"] = hbm_check(eng.n_local, k1_ms / args.steps if args.steps else 0.0)
    if world == 1 and not args.no_cpu:
        try:
            line["cpu_baseline"] = cpu_baseline_block(os.cpu_count() or 1)
        except Exception as exc:  # pragma: no cover
            line["cpu_baseline"] = {"error": str(exc)}
    print(json.dumps(line))
    return 0


def hbm_check(n_local, k1_ms_per_launch):
    """Why `roofline.bound` is not "hbm": K1's ALGORITHMIC bytes per launch (904 B of parameters read + 16 B of results
    written per offspring) over its measured duration, against the measured HBM peak of
    MEASURED_PEAKS.json (driver-written; fallback 6650 GB/s per B200_PROFILING.md)."""
    try:
        alg = D * 4 + 16 * int(n_local)
        peak, src = 6650.0, "of fallback (B200_PROFILING.md)"
        path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(path):
            with open(path) as f:
                peak, src = float(json.load(f)["hbm_gbs"]), "of measured (MEASURED_PEAKS.json hbm_gbs)"
        gbs = alg / (k1_ms_per_launch * 1e-3) / 1e9 if k1_ms_per_launch > 0 else 0.0
        return {"algorithmic_bytes_per_launch": alg, "achieved_gbs": gbs, "peak_gbs": peak, "peak_source": src,
                "frac": gbs / peak, "note": "K1 is five orders of magnitude below the HBM roofline and uses no tensor-core "
                "shaped work (per-offspring weights: GEMV only); the FP32 pipe is the bound that applies"}
    except Exception as exc:  # pragma: no cover
        return {"error": str(exc)}


def measure_fp32_peak(device):
    """FFMA microbenchmark (library test hook) -> TFLOP/s; None if the hook is missing."""
    try:
        import ctypes as C
        from simple_es_b200 import _lib
        lib = _lib.load()
        out = C.c_double(0.0)
        _lib.check(lib.ses_measure_fp32_peak(int(device), C.byref(out)))
        return out.value
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pop", type=int, default=P_DEFAULT)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--exchange", default=os.environ.get("SES_FITNESS_EXCHANGE", "peer"), choices=["peer", "nccl"],
                    help="N > 1: fitness exchange fused into K1 over NVLink peer memory (default) or an NCCL all-gather")
    ap.add_argument("--shard", default="cyclic", choices=["cyclic", "contiguous"],
                    help="N > 1: block-cyclic (default) or contiguous offspring-id ranges per rank")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
