"""Python (float64 / torch-CPU) restatement of the simple-es rollout hot path.

TEST INFRASTRUCTURE ONLY (see oracle/ses_twin.c header).  Three jobs:

1. gym-free / pettingzoo-free environment shims that satisfy the reference's wrapper duck
   type (``reset() -> {agent: {"state": obs}}``, ``step({agent: act}) -> (dict, r, done, info)``,
   ``get_agent_ids()``, ``name``) so that the reference's own ``RolloutWorker`` /
   ``GymEnvModel`` / strategies can be driven in the dev container (oracle/make_golden.py).
2. A port of the reference's per-offspring CPU path (torch policy, Python env, mp.Pool per
   generation) that can travel to the GPU box, where /root/reference does not exist.  It is
   what ``bench.py --impl reference`` and ``cpu_baseline`` time.
3. Readable float64 physics to cross-check the C bit-twin (tests/test_oracle_*.py).

Citations are relative to /root/reference.  The third-party physics (gym CartPole-v1,
PettingZoo MPE simple_spread_v2) is NOT vendored by the reference and not installed here:
it is restated from the published algorithms -- parity unpinned at
that boundary.
"""
import math
import multiprocessing as mp
import time

import numpy as np

# --------------------------------------------------------------------------------------
# CartPole-v1 (gym classic_control/cartpole.py, gym ~0.18-0.21)
# --------------------------------------------------------------------------------------
GRAVITY = 9.8
MASSCART = 1.0
MASSPOLE = 0.1
TOTAL_MASS = MASSPOLE + MASSCART
LENGTH = 0.5
POLEMASS_LENGTH = MASSPOLE * LENGTH
FORCE_MAG = 10.0
TAU = 0.02
THETA_THRESHOLD = 12 * 2 * math.pi / 360
X_THRESHOLD = 2.4


def cartpole_physics(state, action):
    """One Euler step on Python floats, every operation separately rounded, libm sin/cos."""
    x, x_dot, theta, theta_dot = state
    force = FORCE_MAG if action == 1 else -FORCE_MAG
    costheta = math.cos(theta)
    sintheta = math.sin(theta)
    temp = (force + POLEMASS_LENGTH * theta_dot ** 2 * sintheta) / TOTAL_MASS
    thetaacc = (GRAVITY * sintheta - costheta * temp) / (
        LENGTH * (4.0 / 3.0 - MASSPOLE * costheta ** 2 / TOTAL_MASS)
    )
    xacc = temp - POLEMASS_LENGTH * thetaacc * costheta / TOTAL_MASS
    x = x + TAU * x_dot
    x_dot = x_dot + TAU * xacc
    theta = theta + TAU * theta_dot
    theta_dot = theta_dot + TAU * thetaacc
    done = bool(x < -X_THRESHOLD or x > X_THRESHOLD or theta < -THETA_THRESHOLD or theta > THETA_THRESHOLD)
    return (x, x_dot, theta, theta_dot), done


class CartPoleShim:
    """GymWrapper duck type (envs/gym_wrapper.py:7-54) over the restated CartPole-v1.

    ``init_states``: optional [E,4] table cycled by reset() -- the frozen common-random-number
    behaviour the reference shows with process_num > 1; otherwise
    U(-0.05, 0.05)^4 from ``rng``.
    """

    def __init__(self, name="CartPole-v1", max_step=500, pomdp=False, init_states=None, seed=None):
        self.name = name
        self.max_step = max_step
        self.pomdp = pomdp
        self.curr_step = 0
        self.init_states = None if init_states is None else np.asarray(init_states, dtype=np.float64)
        self._reset_count = 0
        self.rng = np.random.RandomState(seed)
        self.state = None

    def _obs(self):
        obs = np.array(self.state, dtype=np.float64)
        if self.pomdp:                       # gym_wrapper.py:73-77 (on the returned copy only)
            obs[1] = 0
            obs[3] = 0
        return obs

    def reset(self):
        self.curr_step = 0                   # gym_wrapper.py:24
        if self.init_states is not None:
            s = self.init_states[self._reset_count % len(self.init_states)]
            self._reset_count += 1
        else:
            s = self.rng.uniform(low=-0.05, high=0.05, size=(4,))
        self.state = tuple(float(v) for v in s)
        return {"0": {"state": self._obs()}}

    def step(self, action):
        self.curr_step += 1                  # gym_wrapper.py:33
        self.state, d = cartpole_physics(self.state, int(action["0"]))
        r = 1.0
        if self.max_step != "None":          # gym_wrapper.py:37-39
            if self.curr_step >= self.max_step or d:
                d = True
        tr = {"state": self._obs(), "reward": r, "done": d, "info": {}}
        return {"0": tr}, r, d, {}

    def get_agent_ids(self):
        return ["0"]

    def close(self):
        pass


# --------------------------------------------------------------------------------------
# MountainCar-v0 and Acrobot-v1 (gym classic_control/mountain_car.py, acrobot.py, gym ~0.18-0.21): any reference
# config can name them through GymWrapper (envs/gym_wrapper.py:8-9 hands the name to gym.make).  Restated from the
# published algorithm; Python floats / numpy float64, libm sin / cos like the originals.
# --------------------------------------------------------------------------------------
def mountaincar_physics(state, action):
    position, velocity = state
    velocity += (action - 1) * 0.001 + math.cos(3 * position) * (-0.0025)
    velocity = float(np.clip(velocity, -0.07, 0.07))
    position += velocity
    position = float(np.clip(position, -1.2, 0.6))
    if position == -1.2 and velocity < 0:
        velocity = 0
    done = bool(position >= 0.5 and velocity >= 0)
    return (position, float(velocity)), -1.0, done


def _acrobot_dsdt(s_augmented):
    m1 = m2 = l1 = 1.0
    lc1 = lc2 = 0.5
    I1 = I2 = 1.0
    g = 9.8
    pi, cos, sin = np.pi, np.cos, np.sin
    a = s_augmented[-1]
    theta1, theta2, dtheta1, dtheta2 = s_augmented[:-1]
    d1 = m1 * lc1 ** 2 + m2 * (l1 ** 2 + lc2 ** 2 + 2 * l1 * lc2 * cos(theta2)) + I1 + I2
    d2 = m2 * (lc2 ** 2 + l1 * lc2 * cos(theta2)) + I2
    phi2 = m2 * lc2 * g * cos(theta1 + theta2 - pi / 2.)
    phi1 = - m2 * l1 * lc2 * dtheta2 ** 2 * sin(theta2) - 2 * m2 * l1 * lc2 * dtheta2 * dtheta1 * sin(theta2) \
        + (m1 * lc1 + m2 * l1) * g * cos(theta1 - pi / 2) + phi2
    ddtheta2 = (a + d2 / d1 * phi1 - m2 * l1 * lc2 * dtheta1 ** 2 * sin(theta2) - phi2) / (m2 * lc2 ** 2 + I2 - d2 ** 2 / d1)
    ddtheta1 = -(d2 * ddtheta2 + phi1) / d1
    return np.array([dtheta1, dtheta2, ddtheta1, ddtheta2, 0.])


def _wrap(x, m, M):
    diff = M - m
    while x > M:
        x = x - diff
    while x < m:
        x = x + diff
    return x


def acrobot_physics(state, action):
    s_augmented = np.append(np.asarray(state, dtype=np.float64), float(action - 1))     # AVAIL_TORQUE = [-1., 0., +1]
    dt = 0.2
    dt2 = dt / 2.0
    y0 = s_augmented
    k1 = _acrobot_dsdt(y0)                                   # rk4(derivs, y0, [0, dt])
    k2 = _acrobot_dsdt(y0 + dt2 * k1)
    k3 = _acrobot_dsdt(y0 + dt2 * k2)
    k4 = _acrobot_dsdt(y0 + dt * k3)
    ns = y0 + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
    ns = [float(v) for v in ns[:4]]
    ns[0] = _wrap(ns[0], -math.pi, math.pi)
    ns[1] = _wrap(ns[1], -math.pi, math.pi)
    ns[2] = min(max(ns[2], -4 * math.pi), 4 * math.pi)
    ns[3] = min(max(ns[3], -9 * math.pi), 9 * math.pi)
    terminal = bool(-np.cos(ns[0]) - np.cos(ns[1] + ns[0]) > 1.)
    return tuple(ns), (-1.0 if not terminal else 0.0), terminal


def pendulum_physics(state, u):
    """gym classic_control/pendulum.py (Pendulum-v0, gym ~0.18): max_speed 8, max_torque 2, dt .05, g 10, m = l = 1.
    `u` is the policy's action widened to a Python float (the real env does `np.clip(u, -2, 2)[0]`, which a 0-d action --
    what GymEnvModel returns for num_action = 1 -- cannot be indexed by: the shim takes the scalar)."""
    th, thdot = state
    g, m, l, dt = 10.0, 1.0, 1.0, 0.05
    u = float(np.clip(u, -2.0, 2.0))
    angle_normalize = ((th + np.pi) % (2 * np.pi)) - np.pi
    costs = angle_normalize ** 2 + .1 * thdot ** 2 + .001 * (u ** 2)
    newthdot = thdot + (-3 * g / (2 * l) * np.sin(th + np.pi) + 3. / (m * l ** 2) * u) * dt
    newth = th + newthdot * dt
    newthdot = float(np.clip(newthdot, -8.0, 8.0))
    return (float(newth), newthdot), float(-costs), False


class ClassicShim:
    """GymWrapper duck type (envs/gym_wrapper.py:7-54) over MountainCar-v0 / Acrobot-v1 / Pendulum-v0; `max_step` is
    min(the config's max_step, gym's TimeLimit) as GymWrapper over a TimeLimit-wrapped env behaves."""
    SPECS = {"MountainCar-v0": (2, 200), "Acrobot-v1": (4, 500), "Pendulum-v0": (2, 200)}

    def __init__(self, name, max_step=None, init_states=None, seed=None):
        self.name = name
        self.state_dim, cap = self.SPECS[name]
        self.max_step = cap if max_step in (None, "None") else min(int(max_step), cap)
        self.curr_step = 0
        self.init_states = None if init_states is None else np.asarray(init_states, dtype=np.float64)
        self._reset_count = 0
        self.rng = np.random.RandomState(seed)
        self.state = None

    def _obs(self):
        s = self.state
        if self.name == "MountainCar-v0":
            return np.array(s, dtype=np.float64)
        if self.name == "Pendulum-v0":
            return np.array([np.cos(s[0]), np.sin(s[0]), s[1]], dtype=np.float64)
        return np.array([np.cos(s[0]), np.sin(s[0]), np.cos(s[1]), np.sin(s[1]), s[2], s[3]], dtype=np.float64)

    def reset(self):
        self.curr_step = 0
        if self.init_states is not None:
            s = self.init_states[self._reset_count % len(self.init_states)]
            self._reset_count += 1
        elif self.name == "MountainCar-v0":
            s = [self.rng.uniform(low=-0.6, high=-0.4), 0.0]
        elif self.name == "Pendulum-v0":
            s = self.rng.uniform(low=[-np.pi, -1.0], high=[np.pi, 1.0])
        else:
            s = self.rng.uniform(low=-0.1, high=0.1, size=(4,))
        self.state = tuple(float(v) for v in s)
        return {"0": {"state": self._obs()}}

    def step(self, action):
        self.curr_step += 1
        if self.name == "Pendulum-v0":
            self.state, r, d = pendulum_physics(self.state, float(action["0"]))      # 0-d float32 array from the tanh head
        else:
            physics = mountaincar_physics if self.name == "MountainCar-v0" else acrobot_physics
            self.state, r, d = physics(self.state, int(action["0"]))
        if self.curr_step >= self.max_step or d:
            d = True
        tr = {"state": self._obs(), "reward": r, "done": d, "info": {}}
        return {"0": tr}, r, d, {}

    def get_agent_ids(self):
        return ["0"]

    def close(self):
        pass


# --------------------------------------------------------------------------------------
# PettingZoo MPE simple_spread_v2, float64 numpy like the original
# --------------------------------------------------------------------------------------
class SimpleSpreadShim:
    """PettingzooWrapper duck type (envs/pettingzoo_wrapper.py:6-64) over a restatement of
    MPE simple_spread_v2 (N agents, N landmarks, local_ratio 0.5, max_cycles 25, discrete)."""

    DT = 0.1
    DAMPING = 0.25
    CONTACT_FORCE = 1e2
    CONTACT_MARGIN = 1e-3
    AGENT_SIZE = 0.15
    SENSITIVITY = 5.0
    MASS = 1.0

    def __init__(self, name="simple_spread", max_step="None", N=2, max_cycles=25, local_ratio=0.5,
                 init_states=None, seed=None):
        self.name = name
        self.max_step = max_step
        self.N = N
        self.max_cycles = max_cycles
        self.local_ratio = local_ratio
        self.curr_step = 0
        self.agents = ["agent_%d" % i for i in range(N)]
        self.init_states = None if init_states is None else np.asarray(init_states, dtype=np.float64)
        self._reset_count = 0
        self.rng = np.random.RandomState(seed)
        self.reset()              # pettingzoo_wrapper.py:20 resets once at construction
        self._reset_count = 0     # ... which must not consume a row of an explicit init table

    # -- world -------------------------------------------------------------------------
    def _observe(self, i):
        other = [self.apos[j] - self.apos[i] for j in range(self.N) if j != i]
        comm = [np.zeros(2) for j in range(self.N) if j != i]
        lm = [self.lpos[k] - self.apos[i] for k in range(self.N)]
        return np.concatenate([self.avel[i], self.apos[i]] + lm + other + comm).astype(np.float32)

    def reset(self):
        self.curr_step = 0
        self.cycles = 0
        if self.init_states is not None:
            s = self.init_states[self._reset_count % len(self.init_states)].reshape(2 * self.N, 2)
            self._reset_count += 1
            self.apos = s[: self.N].copy()
            self.lpos = s[self.N:].copy()
        else:
            self.apos = np.stack([self.rng.uniform(-1, +1, 2) for _ in range(self.N)])
            self.lpos = np.stack([self.rng.uniform(-1, +1, 2) for _ in range(self.N)])
        self.avel = np.zeros((self.N, 2))
        return {a: {"state": self._observe(i)} for i, a in enumerate(self.agents)}

    def _world_step(self, acts):
        N = self.N
        force = np.zeros((N, 2))
        for i, a in enumerate(acts):
            u = np.zeros(2)
            if a == 1:
                u[0] = -1.0
            if a == 2:
                u[0] = +1.0
            if a == 3:
                u[1] = -1.0
            if a == 4:
                u[1] = +1.0
            force[i] = u * self.SENSITIVITY
        for a in range(N):
            for b in range(a + 1, N):
                delta = self.apos[a] - self.apos[b]
                dist = np.sqrt(np.sum(np.square(delta)))
                dist_min = 2 * self.AGENT_SIZE
                k = self.CONTACT_MARGIN
                pen = np.logaddexp(0, -(dist - dist_min) / k) * k
                f = self.CONTACT_FORCE * delta / dist * pen
                force[a] = force[a] + f
                force[b] = force[b] - f
        for i in range(N):
            self.avel[i] = self.avel[i] * (1 - self.DAMPING)
            self.avel[i] += (force[i] / self.MASS) * self.DT
            self.apos[i] += self.avel[i] * self.DT

    def _rewards(self):
        N = self.N
        glob = 0.0
        for k in range(N):
            dists = [np.sqrt(np.sum(np.square(self.apos[a] - self.lpos[k]))) for a in range(N)]
            glob -= min(dists)
        rew = []
        for i in range(N):
            local = 0.0
            for a in range(N):            # 2021 sources: the agent collides with itself as well
                d = np.sqrt(np.sum(np.square(self.apos[a] - self.apos[i])))
                if d < 2 * self.AGENT_SIZE:
                    local -= 1.0
            rew.append(glob * (1 - self.local_ratio) + local * self.local_ratio)
        return rew

    def step(self, action):
        self.curr_step += 1                  # pettingzoo_wrapper.py:34
        acts = [int(action[a]) for a in self.agents]
        self._world_step(acts)               # world steps once after the last agent (:36-42)
        self.cycles += 1
        rew = self._rewards()
        env_done = self.cycles >= self.max_cycles
        ret = {}
        total_r = 0
        for i, a in enumerate(self.agents):
            ret[a] = {"state": self._observe(i), "reward": rew[i], "done": env_done, "info": {}}
            total_r += rew[i]                # pettingzoo_wrapper.py:45-53
        done = env_done
        if self.max_step != "None":          # pettingzoo_wrapper.py:55-57
            if self.curr_step >= self.max_step or done:
                done = True
        return ret, total_r, done, {}

    def get_agent_ids(self):
        return list(self.agents)


# --------------------------------------------------------------------------------------
# flat parameter vector <-> GymEnvModel layout (networks/neural_network.py:12-17)
# --------------------------------------------------------------------------------------
HID = 32


def param_shapes(obs, act, gru):
    shapes = [("fc1.weight", (HID, obs)), ("fc1.bias", (HID,))]
    if gru:
        shapes += [("gru.weight_ih_l0", (3 * HID, HID)), ("gru.weight_hh_l0", (3 * HID, HID)),
                   ("gru.bias_ih_l0", (3 * HID,)), ("gru.bias_hh_l0", (3 * HID,))]
    shapes += [("fc2.weight", (act, HID)), ("fc2.bias", (act,))]
    return shapes


def param_count(obs, act, gru):
    return sum(int(np.prod(s)) for _, s in param_shapes(obs, act, gru))


def flat_to_list(flat, obs, act, gru):
    out, o = [], 0
    for _, s in param_shapes(obs, act, gru):
        n = int(np.prod(s))
        out.append(np.asarray(flat[o:o + n], dtype=np.float32).reshape(s))
        o += n
    return out


def list_to_flat(plist):
    return np.concatenate([np.asarray(p, dtype=np.float32).ravel() for p in plist])


class TorchPolicy:
    """Port of GymEnvModel.forward/reset (networks/neural_network.py:20-40) driven by a flat
    float32 parameter vector; torch CPU ops, so it shows the reference's per-call dispatch cost."""

    def __init__(self, flat, obs, act, gru, discrete=True):
        import torch
        self.torch = torch
        self.gru = gru
        self.discrete = discrete
        p = [torch.from_numpy(np.array(a)) for a in flat_to_list(flat, obs, act, gru)]
        self.w1, self.b1 = p[0], p[1]
        if gru:
            self.cell = torch.nn.GRU(HID, HID)
            with torch.no_grad():
                for dst, src in zip(self.cell.parameters(), p[2:6]):
                    dst.copy_(src)
        self.w2, self.b2 = p[-2], p[-1]
        self.reset()

    def reset(self):
        if self.gru:
            self.h = self.torch.zeros([1, 1, HID], dtype=self.torch.float)

    def __call__(self, x):
        torch = self.torch
        with torch.no_grad():
            x = torch.from_numpy(x).float().unsqueeze(0)
            x = torch.tanh(torch.nn.functional.linear(x, self.w1, self.b1))
            if self.gru:
                x, self.h = self.cell(x, self.h)
                x = torch.tanh(x)
            x = torch.nn.functional.linear(x, self.w2, self.b2)
            if self.discrete:
                x = torch.argmax(torch.nn.functional.softmax(x.squeeze(), dim=0))
            else:
                x = torch.tanh(x.squeeze())
            return x.detach().cpu().numpy()


def rollout_worker(args):
    """Port of RolloutWorker (learning_strategies/evolution/loop.py:108-125)."""
    env, flat, net_cfg, eval_ep_num = args
    obs, act, gru = net_cfg
    agent_ids = env.get_agent_ids()
    models = {k: TorchPolicy(flat, obs, act, gru) for k in agent_ids}
    total_reward = 0
    n_steps = 0
    for _ in range(eval_ep_num):
        states = env.reset()
        done = False
        for m in models.values():
            m.reset()
        while not done:
            actions = {k: m(states[k]["state"][np.newaxis, ...]) for k, m in models.items()}
            states, r, done, _ = env.step(actions)
            total_reward += r
            n_steps += 1
    return total_reward / eval_ep_num, n_steps


# --------------------------------------------------------------------------------------
# strategy ports on flat vectors (offspring_strategies.py, optimizers.py), numpy >= 2 dtypes
# --------------------------------------------------------------------------------------
def argsort_desc_stable(rewards):
    """np.flip(np.argsort(rewards)) with the tie order pinned."""
    return np.flip(np.argsort(np.array(rewards), kind="stable"))


class AdamPort:
    """optimizers.py:7-57 on one flat float32 vector."""

    def __init__(self, D, stepsize, beta1=0.99, beta2=0.999, epsilon=1e-08):
        self.stepsize, self.beta1, self.beta2, self.epsilon = stepsize, beta1, beta2, epsilon
        self.t = 0
        self.m = np.zeros(D, dtype=np.float32)
        self.v = np.zeros(D, dtype=np.float32)

    def lr_t(self):
        return self.stepsize * np.sqrt(1 - self.beta2 ** self.t) / (1 - self.beta1 ** self.t)

    def update(self, theta, g):
        self.t += 1
        a = self.lr_t()
        self.m = self.beta1 * self.m + (1 - self.beta1) * g
        self.v = self.beta2 * self.v + (1 - self.beta2) * (g * g)
        step = -a * self.m / (np.sqrt(self.v) + self.epsilon)
        theta += step                        # float32 += float64 -> rounded once
        return theta


class SGDPort:
    """SGD with momentum in the reference's idiom.  optimizers.py ships Adam only; its header names OpenAI's
    es_distributed/optimizers.py as the source, whose SGD._compute_step is restated here on one flat float32 vector
    (numpy >= 2: the Python floats are weak scalars, every product and sum is float32)."""

    def __init__(self, D, stepsize, momentum=0.9):
        self.stepsize, self.momentum = stepsize, momentum
        self.t = 0
        self.v = np.zeros(D, dtype=np.float32)

    def update(self, theta, g):
        self.t += 1
        self.v = self.momentum * self.v + (1. - self.momentum) * g
        step = -self.stepsize * self.v
        theta += step
        return theta


class StrategyPort:
    """The three offspring strategies on a flat parameter vector, with the reference's
    population layouts, update arithmetic and sigma-decay ordering."""

    def __init__(self, cfg, D):
        self.name = cfg["name"]
        self.D = D
        self.sigma = float(cfg["init_sigma"])
        self.decay = float(cfg["sigma_decay"])
        self.n = int(cfg["offspring_num"])
        self.k = int(cfg.get("elite_num", 1))
        self.lr = float(cfg.get("learning_rate", 0.0))
        self.mu = np.zeros(D, dtype=np.float32)
        self.pop = None
        self.eps = None
        if self.name == "openai_es":
            self.opt = AdamPort(D, self.lr)
        if self.name == "simple_genetic":
            self.elites = np.zeros((self.k, D), dtype=np.float32)

    def _noisy(self, base, scale):
        w = base.copy()
        w += np.random.normal(0, scale, size=w.shape)     # f32 += f64, rounded once
        return w

    def generate(self):
        if self.name == "simple_evolution":               # :165-176
            pop = [self.mu.copy(), self.mu.copy()]
            pop += [self._noisy(self.mu, self.sigma) for _ in range(self.n - 1)]
        elif self.name == "simple_genetic":               # :48-61
            pop = []
            for e in range(self.k):
                pop.append(self.elites[e].copy())
                pop += [self._noisy(self.elites[e], self.sigma) for _ in range(self.n // self.k - 1)]
        else:                                             # openai_es :299-328
            pop, eps = [self.mu.copy()], [self.mu.copy()]
            for _ in range(self.n - 1):
                e = np.random.normal(size=self.D)
                w = self.mu.copy(); w += e * self.sigma
                s = self.mu.copy(); s += e                # quirk Q1: stored "epsilon" is mu+eps
                pop.append(w); eps.append(s)
            self.eps = np.stack(eps)
        self.pop = np.stack(pop)
        return self.pop

    def evaluate(self, rewards):
        rewards = list(rewards)
        order = argsort_desc_stable(rewards)
        best = max(rewards)
        if self.name == "simple_evolution":               # :234-258
            acc = self.pop[order[0]].copy()
            for e in order[1:self.k]:
                acc += self.pop[e]
            acc /= self.k
            self.mu = acc
            self.sigma *= self.decay
        elif self.name == "simple_genetic":               # :112-124
            self.elites = self.pop[order[:self.k]].copy()
        else:                                             # :380-418
            P = len(rewards)
            shaped = np.zeros(P)
            for idx in reversed(range(P)):
                shaped[order[idx]] = ((P - 1 - idx) / (P - 1)) - 0.5
            shaped = (shaped - shaped.mean()) / shaped.std()
            g = np.zeros(self.D, dtype=np.float32)
            for j in range(P):
                g += self.eps[j] * shaped[j]
            g *= -1.0 * (self.lr / (P * self.sigma))
            self.mu = self.opt.update(self.mu, g)
            self.sigma *= self.decay
        pop = self.generate()
        if self.name == "simple_genetic":
            self.sigma *= self.decay                      # genetic decays AFTER regenerating (:117-124)
        return pop, best, self.sigma


def es_loop_port(env, net_cfg, strategy_cfg, generation_num, process_num, eval_ep_num, seed=0, verbose=False, state=None):
    """Port of ESLoop.run (loop.py:52-104): a fresh mp.Pool every generation, one task per
    offspring carrying (env, weights, eval_ep_num).  Returns per-generation records.
    `state` (optional dict: mu, sigma and, for openai_es, m / v / t) starts the run from a trained
    strategy state instead of the reference's all-zero network (loop.py:31) -- used by bench.py to time
    the reference path in the regime the GPU arm is timed in."""
    np.random.seed(seed)
    D = param_count(*net_cfg)
    strat = StrategyPort(strategy_cfg, D)
    if state is not None:
        strat.mu = np.array(state["mu"], dtype=np.float32)
        strat.sigma = float(state.get("sigma", strat.sigma))
        if strat.name == "openai_es" and "m" in state:
            strat.opt.m = np.array(state["m"], dtype=np.float32)
            strat.opt.v = np.array(state["v"], dtype=np.float32)
            strat.opt.t = int(state["t"])
    pop = strat.generate()
    out = []
    for g in range(generation_num):
        t0 = time.time()
        p = mp.Pool(process_num) if process_num > 1 else None
        args = [(env, w, net_cfg, eval_ep_num) for w in pop]
        t1 = time.time()
        res = p.map(rollout_worker, args) if p is not None else [rollout_worker(a) for a in args]
        if p is not None:
            p.close()
        t2 = time.time()
        rewards = [r for r, _ in res]
        steps = sum(s for _, s in res)
        pop, best, sigma = strat.evaluate(rewards)
        t3 = time.time()
        rec = dict(gen=g, best=best, sigma=sigma, env_steps=steps, time=t3 - t0, rollout_t=t2 - t1, eval_t=t3 - t2)
        if verbose:
            print(rec)
        out.append(rec)
    return out
