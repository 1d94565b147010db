"""Generate tests/golden/*.npz by running the UNMODIFIED reference classes (dev container only).

    python oracle/make_golden.py

Every array written here is an output (or an input) of the reference's own code
(/root/reference: GymEnvModel, RolloutWorker, simple_evolution, simple_genetic, openai_es, Adam)
driven through the duck-typed env shims of oracle/pyref.py.  np.argsort is pinned to
kind="stable" while evaluate() runs.  The fixtures are committed; this
script is committed so they can be regenerated.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import pyref, ref_bridge  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def flat_of(model):
    return np.concatenate([p.ravel() for p in model.get_param_list()]).astype(np.float32)


def set_flat(model, flat, obs, act, gru):
    model.apply_param(pyref.flat_to_list(flat, obs, act, gru))


class TracingCartPole(pyref.CartPoleShim):
    """CartPoleShim that records (state after step, action) for every step."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.log = []

    def reset(self):
        self.log.append(("reset",))
        return super().reset()

    def step(self, action):
        out = super().step(action)
        self.log.append((int(action["0"]), tuple(self.state)))
        return out


def golden_policy(ref, gru, n_models, n_steps, seed):
    rng = np.random.RandomState(seed)
    obs_dim, act = 4, 2
    D = pyref.param_count(obs_dim, act, gru)
    W = (rng.normal(0, 2.0 if not gru else 0.7, size=(n_models, D))).astype(np.float32)
    O = rng.uniform(-1, 1, size=(n_models, n_steps, obs_dim)) * np.array([2.4, 3.0, 0.21, 3.0])
    A = np.zeros((n_models, n_steps), dtype=np.int64)
    Z = np.zeros((n_models, n_steps, act), dtype=np.float32)
    for i in range(n_models):
        model = ref.GymEnvModel(obs_dim, act, True, gru)
        set_flat(model, W[i], obs_dim, act, gru)
        cap = []
        hook = model.fc2.register_forward_hook(lambda m, inp, out: cap.append(out.detach().numpy().ravel().copy()))
        model.reset()
        for t in range(n_steps):
            A[i, t] = int(model(O[i, t][np.newaxis, ...]))
            Z[i, t] = cap[-1]
        hook.remove()
    return dict(W=W, obs=O, actions=A, logits=Z, gru=np.int32(gru))


def golden_rollout(ref, gru, pomdp, P, E, seed, sigma, n_trace):
    """Reference RolloutWorker + GymEnvModel over the CartPole shim with a fixed [E,4] table of
    initial states shared by every offspring (pool semantics)."""
    rng = np.random.RandomState(seed)
    obs_dim, act = 4, 2
    D = pyref.param_count(obs_dim, act, gru)
    init = rng.uniform(-0.05, 0.05, size=(E, 4))
    W = rng.normal(0, sigma, size=(P, D)).astype(np.float32)
    W[0] = 0.0
    fitness = np.zeros(P)
    logs = []
    for i in range(P):
        model = ref.GymEnvModel(obs_dim, act, True, gru)
        set_flat(model, W[i], obs_dim, act, gru)
        env = TracingCartPole(max_step=500, pomdp=pomdp, init_states=init)
        fitness[i] = ref.RolloutWorker((env, {"0": model}, E))
        ep0 = []
        for rec in env.log[1:]:
            if rec[0] == "reset":
                break
            ep0.append(rec)
        logs.append(ep0)
    # trace the offspring whose first episode is longest (first 200 steps of episode 0)
    trace_ids = np.argsort([-len(l) for l in logs], kind="stable")[:n_trace].astype(np.int32)
    traces = np.full((n_trace, 200, 4), np.nan)
    tr_actions = np.full((n_trace, 200), -1, dtype=np.int32)
    for j, i in enumerate(trace_ids):
        for t, rec in enumerate(logs[i][:200]):
            tr_actions[j, t] = rec[0]
            traces[j, t] = rec[1]
    return dict(W=W, init=init, fitness=fitness, trace_ids=trace_ids, traces=traces, trace_actions=tr_actions,
                gru=np.int32(gru), pomdp=np.int32(pomdp), E=np.int32(E), max_step=np.int32(500))


def large_population(seed, P, D=226):
    """The weights of the large CartPole golden set, regenerated from the seed by the tests (numpy's legacy RandomState stream is
    frozen): the first half random policies (sigma 2, episodes of ~10-60 steps), the second half perturbations (sigma 0.3) of a
    hand-built balancing parent (episodes from a few steps up to the 500-step limit)."""
    rng = np.random.RandomState(seed)
    init = rng.uniform(-0.05, 0.05, size=(5, 4))
    W = rng.normal(0, 2.0, size=(P, D)).astype(np.float32)
    base = np.zeros(D, np.float32)
    base[:128].reshape(32, 4)[0] = [0.0, 0.5, 10.0, 3.0]
    base[160:224].reshape(2, 32)[1, 0] = 5.0
    base[160:224].reshape(2, 32)[0, 0] = -5.0
    W[P // 2:] = base + (W[P // 2:] * np.float32(0.15))
    return init, W


def golden_rollout_large(ref, P, seed):
    """>= 4096 offspring so that 'at least 99.9 % of the returns equal the reference's' means something.
    Only the returns are stored (32 kB); tests rebuild the weights with large_population()."""
    init, W = large_population(seed, P)
    E = 5
    fitness = np.zeros(P)
    for i in range(P):
        model = ref.GymEnvModel(4, 2, True, False)
        set_flat(model, W[i], 4, 2, False)
        env = pyref.CartPoleShim(max_step=500, init_states=init)
        fitness[i] = ref.RolloutWorker((env, {"0": model}, E))
    import zlib
    return dict(seed=np.int64(seed), P=np.int32(P), E=np.int32(E), fitness=fitness, init=init,
                w_crc32=np.uint32(zlib.crc32(np.ascontiguousarray(W).tobytes())))


class TracingSpread(pyref.SimpleSpreadShim):
    def __init__(self, *a, **kw):
        self.log = []
        super().__init__(*a, **kw)

    def reset(self):
        self.log.append(("reset",))
        return super().reset()

    def step(self, action):
        out = super().step(action)
        self.log.append(([int(action[a]) for a in self.agents], self.apos.copy().ravel(), self.avel.copy().ravel(), out[1]))
        return out


def golden_spread(ref, N, P, E, seed, sigma, n_trace, gru=False):
    """Reference RolloutWorker + one GymEnvModel copy per agent (the reference's own utils.wrap_agentid: a deepcopy per agent id,
    so with gru=True every agent has its own hidden state) over the simple_spread restatement, fixed [E, 4N] initial positions
    shared by every offspring."""
    rng = np.random.RandomState(seed)
    obs_dim, act = 6 * N, 5
    D = pyref.param_count(obs_dim, act, gru)
    init = rng.uniform(-1, 1, size=(E, 4 * N))
    W = rng.normal(0, sigma, size=(P, D)).astype(np.float32)
    W[0] = 0.0
    fitness = np.zeros(P)
    traces = np.full((n_trace, 25, 4 * N), np.nan)
    tr_actions = np.full((n_trace, 25, N), -1, dtype=np.int32)
    tr_rewards = np.full((n_trace, 25), np.nan)
    for i in range(P):
        model = ref.GymEnvModel(obs_dim, act, True, gru)
        set_flat(model, W[i], obs_dim, act, gru)
        env = TracingSpread(N=N, init_states=init)
        env.log = []
        group = ref.wrap_agentid(env.get_agent_ids(), model)           # learning_strategies/evolution/utils.py:4-8
        fitness[i] = ref.RolloutWorker((env, group, E))
        if i < n_trace:
            for t, rec in enumerate(env.log[1:26]):
                tr_actions[i, t] = rec[0]
                traces[i, t] = np.concatenate([rec[1], rec[2]])
                tr_rewards[i, t] = rec[3]
    return dict(W=W, init=init, fitness=fitness, traces=traces, trace_actions=tr_actions, trace_rewards=tr_rewards,
                N=np.int32(N), E=np.int32(E), gru=np.int32(gru))


def pyref_continuous():
    return ("Pendulum-v0",)


class TracingClassic(pyref.ClassicShim):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.log = []

    def reset(self):
        self.log.append(("reset",))
        return super().reset()

    def step(self, action):
        out = super().step(action)
        a = action["0"]
        self.log.append((float(a) if self.name in pyref_continuous() else int(a), tuple(self.state), out[1]))
        return out


def golden_classic(ref, env_name, P, E, seed, sigma, n_trace, gru=False):
    """Reference RolloutWorker + GymEnvModel over MountainCar-v0 / Acrobot-v1 (oracle/pyref.py::ClassicShim), a fixed
    [E, state_dim] table of initial states shared by every offspring (pool semantics)."""
    rng = np.random.RandomState(seed)
    sd, cap = pyref.ClassicShim.SPECS[env_name]
    obs_dim, act = {"MountainCar-v0": (2, 3), "Acrobot-v1": (6, 3), "Pendulum-v0": (3, 1)}[env_name]
    discrete = env_name not in pyref_continuous()          # Pendulum: GymEnvModel(discrete_action=False), the tanh head
    D = pyref.param_count(obs_dim, act, gru)
    if env_name == "MountainCar-v0":
        init = np.stack([rng.uniform(-0.6, -0.4, size=E), np.zeros(E)], axis=1)
    elif env_name == "Pendulum-v0":
        init = rng.uniform([-np.pi, -1.0], [np.pi, 1.0], size=(E, 2))
    else:
        init = rng.uniform(-0.1, 0.1, size=(E, 4))
    W = rng.normal(0, sigma, size=(P, D)).astype(np.float32)
    W[0] = 0.0
    fitness = np.zeros(P)
    steps = np.zeros(P, dtype=np.int64)
    logs = []
    for i in range(P):
        model = ref.GymEnvModel(obs_dim, act, discrete, gru)
        set_flat(model, W[i], obs_dim, act, gru)
        env = TracingClassic(env_name, max_step=cap, init_states=init)
        fitness[i] = ref.RolloutWorker((env, {"0": model}, E))
        steps[i] = sum(1 for rec in env.log if rec[0] != "reset")
        ep0 = []
        for rec in env.log[1:]:
            if rec[0] == "reset":
                break
            ep0.append(rec)
        logs.append(ep0)
    trace_ids = np.arange(n_trace, dtype=np.int32) + 1           # offspring 1..n_trace (0 is the all-zero policy)
    traces = np.full((n_trace, 200, sd), np.nan)
    tr_actions = np.full((n_trace, 200), -1, dtype=np.int32) if discrete else np.full((n_trace, 200), np.nan, dtype=np.float32)
    for j, i in enumerate(trace_ids):
        for t, rec in enumerate(logs[i][:200]):
            tr_actions[j, t] = rec[0]
            traces[j, t] = rec[1]
    return dict(W=W, init=init, fitness=fitness, steps=steps, trace_ids=trace_ids, traces=traces, trace_actions=tr_actions,
                E=np.int32(E), max_step=np.int32(cap))


def reward_vectors(P, rng):
    """Three synthetic reward vectors: tie-free floats, CartPole-like tie-heavy k/5, mixed."""
    r0 = rng.uniform(8, 500, size=P)
    r1 = rng.randint(40, 60, size=P) / 5.0
    r1[rng.randint(0, P, size=max(2, P // 4))] = 500.0
    r2 = np.round(rng.uniform(8, 30, size=P))
    return [r0, r1, r2]


def alias_reward_vectors(P, rng):
    """simple_evolution's object aliasing (offspring_strategies.py:165-176,232-250: population slots 0 and 1 are `mu_model` and
    `elite_models[0]`, the SAME module at generation 0 and whenever slot 0 or 1 won the previous generation; the elite sum runs in
    place on the winner's storage).  Four generations that put the two aliased slots into the elite set in every position:
    first and second (generation 0: the all-zero network added to itself; generation 1: non-zero weights, x += x), in the middle
    of the elites behind another winner, and inside a larger group of tied rewards."""
    base = lambda: rng.uniform(8, 400, size=P)
    r0 = base(); r0[0] = r0[1] = 500.0
    r1 = base(); r1[0] = r1[1] = 500.0
    r2 = base(); r2[7] = 500.0; r2[0] = r2[1] = 450.0
    r3 = base(); r3[[0, 1, 5, 9]] = 500.0
    return [r0, r1, r2, r3]


def golden_strategy(ref, name, seed, rewards_fn=None):
    obs_dim, act, gru = 4, 2, False
    D = pyref.param_count(obs_dim, act, gru)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if name == "simple_evolution":
        cfg = dict(init_sigma=2.0, sigma_decay=0.99, elite_num=5, offspring_num=24)
        strat = ref.simple_evolution(cfg["init_sigma"], cfg["sigma_decay"], cfg["elite_num"], cfg["offspring_num"])
    elif name == "simple_genetic":
        cfg = dict(init_sigma=1.0, sigma_decay=0.98, elite_num=4, offspring_num=26)
        strat = ref.simple_genetic(cfg["init_sigma"], cfg["sigma_decay"], cfg["elite_num"], cfg["offspring_num"])
    else:
        cfg = dict(init_sigma=0.2, sigma_decay=0.999, learning_rate=0.1, offspring_num=40)
        strat = ref.openai_es(cfg["init_sigma"], cfg["sigma_decay"], cfg["learning_rate"], cfg["offspring_num"])
    net = ref.GymEnvModel(obs_dim, act, True, gru)
    net.zero_init()
    group = strat.init_offspring(net, ["0"])
    out = {"cfg_" + k: np.float64(v) for k, v in cfg.items()}
    rng = np.random.RandomState(seed + 1)
    P = len(group)
    rews = (rewards_fn or reward_vectors)(P, rng)
    out["P"] = np.int32(P)
    out["generations"] = np.int32(len(rews))
    for g, rewards in enumerate(rews):
        pop = np.stack([flat_of(off["0"]) for off in group])
        out["pop_%d" % g] = pop
        out["rewards_%d" % g] = rewards
        out["sigma_before_%d" % g] = np.float64(strat.curr_sigma)
        if name == "openai_es":
            out["eps_%d" % g] = np.stack([flat_of(e) for e in strat.epsilons])
            out["mu_before_%d" % g] = flat_of(strat.mu_model)
        sink = []
        with ref_bridge.stable_argsort(ref.strategies), ref_bridge.capture_locals("evaluate", sink):
            group, best, sigma = strat.evaluate(list(rewards))
        loc = sink[-1]
        out["best_%d" % g] = np.float64(best)
        out["sigma_after_%d" % g] = np.float64(sigma)
        if name == "openai_es":
            out["order_%d" % g] = np.asarray(loc["offspring_rank_id"], dtype=np.int64)
            out["shaped_%d" % g] = np.asarray(loc["reward_array"], dtype=np.float64)
            out["update_factor_%d" % g] = np.float64(loc["update_factor"])
            out["grad_%d" % g] = np.concatenate([p.ravel() for p in loc["grad_param_list"]]).astype(np.float32)
            out["mu_after_%d" % g] = flat_of(strat.mu_model)
            out["adam_m_%d" % g] = np.concatenate([p.ravel() for p in strat.optimizer.m]).astype(np.float32)
            out["adam_v_%d" % g] = np.concatenate([p.ravel() for p in strat.optimizer.v]).astype(np.float32)
            out["adam_t_%d" % g] = np.int64(strat.optimizer.t)
        else:
            out["elite_ids_%d" % g] = np.asarray(loc["elite_ids"], dtype=np.int64)
            if name == "simple_evolution":
                out["mu_after_%d" % g] = flat_of(strat.mu_model)
            else:
                out["elites_after_%d" % g] = np.stack([flat_of(m) for m in strat.elite_models])
    out["pop_final"] = np.stack([flat_of(off["0"]) for off in group])
    # unpinned argsort on the tie-free vector must agree with the pinned order
    out["order_unpinned_0"] = np.flip(np.argsort(np.array(list(rews[0])))).astype(np.int64)
    return out


def main():
    ref = ref_bridge.load()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    jobs = {
        "policy_mlp": lambda: golden_policy(ref, False, 48, 32, 11),
        "policy_gru": lambda: golden_policy(ref, True, 12, 40, 12),
        "rollout_cartpole_mlp": lambda: golden_rollout(ref, False, False, 256, 5, 21, 2.0, 6),
        "rollout_cartpole_mlp_4096": lambda: golden_rollout_large(ref, 4096, 71),
        "rollout_cartpole_gru_pomdp": lambda: golden_rollout(ref, True, True, 24, 3, 22, 0.7, 3),
        "rollout_spread_n2": lambda: golden_spread(ref, 2, 128, 5, 41, 1.0, 4),
        "rollout_spread_n3": lambda: golden_spread(ref, 3, 48, 3, 42, 1.0, 2),
        "rollout_mountaincar": lambda: golden_classic(ref, "MountainCar-v0", 96, 3, 51, 3.0, 4),
        "rollout_acrobot": lambda: golden_classic(ref, "Acrobot-v1", 64, 3, 52, 2.0, 4),
        "rollout_pendulum": lambda: golden_classic(ref, "Pendulum-v0", 96, 3, 53, 1.0, 4),
        # the recurrent policy (`gru: True`) beyond CartPole: one hidden state per agent copy (utils.wrap_agentid)
        "rollout_spread_n2_gru": lambda: golden_spread(ref, 2, 32, 3, 61, 0.5, 4, gru=True),
        "rollout_spread_n3_gru": lambda: golden_spread(ref, 3, 16, 2, 62, 0.5, 2, gru=True),
        "rollout_mountaincar_gru": lambda: golden_classic(ref, "MountainCar-v0", 24, 2, 63, 1.0, 4, gru=True),
        "rollout_pendulum_gru": lambda: golden_classic(ref, "Pendulum-v0", 24, 2, 64, 0.5, 4, gru=True),
        "rollout_acrobot_gru": lambda: golden_classic(ref, "Acrobot-v1", 16, 2, 65, 0.7, 2, gru=True),
        "strategy_simple_evolution": lambda: golden_strategy(ref, "simple_evolution", 31),
        "strategy_simple_evolution_alias": lambda: golden_strategy(ref, "simple_evolution", 34, alias_reward_vectors),
        "strategy_simple_genetic": lambda: golden_strategy(ref, "simple_genetic", 32),
        "strategy_openai_es": lambda: golden_strategy(ref, "openai_es", 33),
    }
    only = sys.argv[1:]
    for name, fn in jobs.items():
        if only and name not in only:
            continue
        data = fn()
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **data)
        print("wrote", path, {k: getattr(v, "shape", None) for k, v in data.items() if hasattr(v, "shape") and v.ndim > 0})


if __name__ == "__main__":
    main()
