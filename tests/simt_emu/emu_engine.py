"""numpy front end of libses_simt_emu.so (the engine's CUDA sources compiled for the host on the SIMT emulator,
tests/simt_emu/build.py).  TEST INFRASTRUCTURE ONLY: mirrors simple_es_b200.engine.RolloutEngine call for call so
that the CPU tests read like the GPU parity tests, with numpy arrays standing in for device memory."""
import ctypes as C
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from simple_es_b200 import _lib as product_lib  # noqa: E402  (signatures and the ses_config struct only)
from simple_es_b200.engine import CONTINUOUS_ENVS, ENV_IDS, ENV_SPECS, owned_ids  # noqa: E402  (pure-Python host logic)

from . import build as emu_build  # noqa: E402

_lib = None


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(emu_build.build())
        for name, (res, args) in dict(product_lib.SYMBOLS, **product_lib.TEST_SYMBOLS).items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _arr(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


class EmuEngine:
    def __init__(self, env_name="CartPole-v1", obs_dim=4, act_dim=2, gru=False, pomdp=False, max_step=500, eval_ep_num=5,
                 population=256, group=256, n_head=1, n_parents=1, seed=0, init_mode="shared", n_agents=2, id_begin=0,
                 id_end=None, antithetic=False, shard=None):
        self.lib = load()
        self.P = int(population)
        self.id_begin, self.id_end = int(id_begin), self.P if id_end is None else int(id_end)
        self.shard = None if shard is None else tuple(int(x) for x in shard)
        self.n_local = int(owned_ids(self.P, *self.shard).size) if self.shard else self.id_end - self.id_begin
        self.E = int(eval_ep_num)
        self.n_agents = int(n_agents) if env_name == "simple_spread" else 1
        _, cap, sdim, _ = ENV_SPECS[env_name]
        ms = 0 if max_step in (None, "None") else int(max_step)
        self.max_step = min(ms, cap) if ms > 0 else cap
        self.state_dim = 4 * self.n_agents if sdim is None else sdim
        self.D = self.lib.ses_param_count(obs_dim, act_dim, int(bool(gru)))
        self.cfg = product_lib.ses_config(
            env=ENV_IDS[env_name], obs_dim=obs_dim, act_dim=act_dim, gru=int(bool(gru)), pomdp=int(bool(pomdp)),
            n_agents=self.n_agents, max_step=self.max_step, eval_ep_num=self.E, population=self.P, group=int(group),
            n_head=int(n_head), n_parents=int(n_parents), seed=int(seed) & 0xFFFFFFFF,
            init_mode={"shared": 0, "fresh": 1}[init_mode], id_begin=self.id_begin, id_end=self.id_end, device=0,
            antithetic=int(bool(antithetic)), shard_block=self.shard[2] if self.shard else 0,
            shard_rank=self.shard[0] if self.shard else 0, shard_world=self.shard[1] if self.shard else 0,
            continuous_action=int(env_name in CONTINUOUS_ENVS))
        h = C.c_void_p()
        self._check(self.lib.ses_create(C.byref(self.cfg), C.byref(h)))
        self._h = h
        if env_name != "simple_spread" and env_name not in CONTINUOUS_ENVS:
            self.key_bits, self.key_scale = int(self.E * self.max_step).bit_length(), float(self.E)
        else:
            self.key_bits, self.key_scale = 0, 1.0

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("simt_emu: " + self.lib.ses_last_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            self.lib.ses_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    @property
    def launches(self):
        return int(self.lib.ses_launch_count(self._h))

    def rollout(self, generation, sigma, parents, w_override=None, init_states=None, n_trace=0, fitness=None, steps=None):
        parents = _arr(parents, np.float32)
        w_override = _arr(w_override, np.float32)
        init_states = _arr(init_states, np.float64)
        fitness = np.full(self.P, np.nan) if fitness is None else fitness
        steps = np.full(self.P, -1, dtype=np.int64) if steps is None else steps
        trace = actions = None
        if n_trace > 0:
            n_trace = min(n_trace, self.n_local)
            trace = np.full((n_trace, 200, self.state_dim), np.nan)
            actions = np.full((n_trace, 200, self.n_agents), -1, dtype=np.int32)
        self._check(self.lib.ses_rollout(self._h, int(generation), float(sigma), _p(parents), _p(w_override), _p(init_states),
                                         _p(fitness), _p(steps), _p(trace), _p(actions), int(n_trace), None))
        return (fitness, steps, trace, actions) if n_trace > 0 else (fitness, steps)

    def rank_desc(self, fitness, shaped=False, full_key=False):
        fitness = _arr(fitness, np.float64)
        n = fitness.size
        order = np.full(n, -1, dtype=np.int32)
        sh = np.full(n, np.nan) if shaped else None
        kb, ks = (0, 1.0) if full_key else (self.key_bits, self.key_scale)
        self._check(self.lib.ses_rank_desc(self._h, _p(fitness), n, kb, ks, _p(order), _p(sh), None))
        return (order, sh) if shaped else order

    @staticmethod
    def adam_a(lr, t, beta1=0.99, beta2=0.999):
        return lr * math.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)

    def update_openai(self, generation, sigma, lr, t, shaped, mu, m, v, eps_override=None, beta1=0.99, beta2=0.999, eps=1e-8):
        """mu / m / v: float32 arrays updated in place; returns the scaled gradient."""
        for a in (mu, m, v):
            assert a.dtype == np.float32 and a.flags.c_contiguous and a.size == self.D
        grad = np.full(self.D, np.nan, dtype=np.float32)
        shaped = _arr(shaped, np.float64)
        eps_override = _arr(eps_override, np.float32)
        uf = -1.0 * (lr / (self.P * sigma))
        self._check(self.lib.ses_update_openai(self._h, int(generation), _p(shaped), _p(eps_override), uf,
                                               self.adam_a(lr, t, beta1, beta2), beta1, beta2, eps, _p(mu), _p(m), _p(v),
                                               _p(grad), None))
        return grad

    def update_openai_sgd(self, generation, sigma, lr, shaped, mu, v, momentum=0.9, eps_override=None):
        grad = np.full(self.D, np.nan, dtype=np.float32)
        uf = -1.0 * (lr / (self.P * sigma))
        self._check(self.lib.ses_update_openai_sgd(self._h, int(generation), _p(_arr(shaped, np.float64)), _p(_arr(eps_override, np.float32)),
                                                   uf, float(lr), float(momentum), _p(mu), _p(v), _p(grad), None))
        return grad

    def materialize(self, generation, sigma, parents, ids, w_override=None):
        ids = _arr(ids, np.int32)
        out = np.full((ids.size, self.D), np.nan, dtype=np.float32)
        self._check(self.lib.ses_materialize(self._h, int(generation), float(sigma), _p(_arr(parents, np.float32)),
                                             _p(_arr(w_override, np.float32)), _p(ids), ids.size, _p(out), None))
        return out

    def elite_mean(self, generation, sigma, parents, order, k, w_override=None):
        out = np.full(self.D, np.nan, dtype=np.float32)
        self._check(self.lib.ses_update_elite_mean(self._h, int(generation), float(sigma), _p(_arr(parents, np.float32)),
                                                   _p(_arr(w_override, np.float32)), _p(_arr(order, np.int32)), int(k), _p(out), None))
        return out

    def generation_openai_host(self, generation, sigma, lr, t, mu, m, v, fitness):
        total = np.zeros(1, dtype=np.int64)
        self._check(self.lib.ses_generation_openai_host(self._h, int(generation), float(sigma), float(lr), int(t), _p(mu), _p(m),
                                                        _p(v), _p(fitness), _p(total), None))
        return int(total[0])

    # ------------------------------------------------------------------ peer exchange (emulated ranks = threads of this process)
    def peer_export(self):
        mine = C.create_string_buffer(64)
        self._check(self.lib.ses_peer_export(self._h, mine))
        return mine.raw

    def peer_attach(self, handles, rank, world):
        """-> the two [P] float64 exchange buffers (views of library-owned memory), by generation parity."""
        blob = C.create_string_buffer(b"".join(handles), 64 * world)
        self._check(self.lib.ses_peer_attach(self._h, blob, int(rank), int(world)))
        bufs = []
        for parity in (0, 1):
            ptr = C.c_void_p()
            self._check(self.lib.ses_peer_fitness_ptr(self._h, parity, C.byref(ptr)))
            bufs.append(np.ctypeslib.as_array((C.c_double * self.P).from_address(ptr.value)))
        return bufs

    def peer_barrier(self):
        self._check(self.lib.ses_peer_barrier(self._h, None))

    def peer_check(self):
        self._check(self.lib.ses_peer_check(self._h))

    def generation_evolution_host(self, generation, sigma, elite_num, mu, fitness):
        total = np.zeros(1, dtype=np.int64)
        self._check(self.lib.ses_generation_evolution_host(self._h, int(generation), float(sigma), int(elite_num), _p(mu), _p(fitness),
                                                           _p(total), None))
        return int(total[0])

    def generation_genetic_host(self, generation, sigma, elites, fitness):
        total = np.zeros(1, dtype=np.int64)
        self._check(self.lib.ses_generation_genetic_host(self._h, int(generation), float(sigma), _p(elites), _p(fitness), _p(total), None))
        return int(total[0])

    def test_math(self, kind, x):
        kinds = {"tanh": 0, "sigmoid": 1, "ln": 2, "sin2pi": 3, "cos2pi": 4, "sin64": 5, "cos64": 6, "tanh_fast": 7,
                 "sin64_full": 8, "cos64_full": 9}
        x = np.ascontiguousarray(x)
        out = np.empty_like(x)
        self._check(self.lib.ses_test_math(kinds[kind], _p(x), _p(out), x.size, None))
        return out

    def test_k1_geometry(self):
        out = (C.c_int32 * 8)()
        self._check(self.lib.ses_test_k1_geometry(self._h, out))
        return dict(zip(("grid", "lanes", "tail_start", "sparse_rank", "sparse_quota", "ctas_per_sm", "resident_warps", "reserved"), list(out)))

    def test_normals(self, generation, idx):
        out = np.full(self.D, np.nan, dtype=np.float32)
        self._check(self.lib.ses_test_normals(self._h, int(generation), int(idx), _p(out), None))
        return out

    def test_ddiv_fast(self, n):
        bad = C.c_uint64(0)
        self._check(self.lib.ses_test_ddiv_fast(int(n), C.byref(bad)))
        return int(bad.value)

    def test_tanh_x2_exhaustive(self, newton, lo, hi):
        bad = C.c_uint64(0)
        self._check(self.lib.ses_test_tanh_x2_exhaustive(int(bool(newton)), float(lo), float(hi), C.byref(bad)))
        return int(bad.value)
