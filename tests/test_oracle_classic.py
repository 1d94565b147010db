"""MountainCar-v0 / Acrobot-v1: the C bit-twin (oracle/ses_twin_classic.c) against numpy's sin / cos, against the
independent float64 Python restatement (oracle/pyref.py, libm trigonometry) and against the reference-driven golden
vectors (the reference's own RolloutWorker + GymEnvModel over oracle/pyref.py::ClassicShim)."""
import math

import numpy as np
import pytest

ENVS = ["MountainCar-v0", "Acrobot-v1"]
GOLD = {"MountainCar-v0": "rollout_mountaincar", "Acrobot-v1": "rollout_acrobot"}


def test_sincos_full_within_one_ulp_of_libm(twin):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-100, 100, 400_000), rng.uniform(-1e-3, 1e-3, 1000), np.arange(-40, 41) * (math.pi / 4),
                        [0.0, -0.0, 3 * -1.2, 3 * 0.6, math.pi, -math.pi, math.pi / 2]])
    s, c = twin.sincos_full(x)
    assert np.all(np.abs(s - np.sin(x)) <= np.spacing(np.abs(np.sin(x)))) and np.all(np.abs(c - np.cos(x)) <= np.spacing(np.abs(np.cos(x))))
    assert np.all(s * s + c * c - 1.0 < 4e-16)


def test_mountaincar_known_answers(twin):
    # one step from (-0.5, 0) pushing right: v = 0.001 - 0.0025 cos(-1.5); x = -0.5 + v
    st, r, done = twin.classic_step("MountainCar-v0", [-0.5, 0.0], 2)
    v = 0.001 + math.cos(-1.5) * (-0.0025)
    assert abs(st[1] - v) < 1e-18 and abs(st[0] - (-0.5 + v)) < 1e-16 and r == -1.0 and not done
    # inelastic left wall, speed clip, goal
    st, r, done = twin.classic_step("MountainCar-v0", [-1.199, -0.07], 0)
    assert st[0] == -1.2 and st[1] == 0.0 and not done
    st, _, _ = twin.classic_step("MountainCar-v0", [0.0, 0.0699], 2)
    assert st[1] <= 0.07
    st, r, done = twin.classic_step("MountainCar-v0", [0.49, 0.05], 2)
    assert done and r == -1.0 and st[0] >= 0.5
    # the all-zero policy always answers action 0 (push left): never reaches the goal, 200 steps, return -200
    f, n = twin.rollout_classic("MountainCar-v0", np.zeros(195, np.float32), E=2)
    assert f == -200.0 and n == 400


def test_acrobot_known_answers(twin):
    # hanging at rest with zero torque is an equilibrium of the dynamics
    st, r, done = twin.classic_step("Acrobot-v1", [0.0, 0.0, 0.0, 0.0], 1)
    assert np.all(np.abs(st) < 1e-15) and r == -1.0 and not done
    # upright (theta1 = pi) is above the bar: terminal, reward 0
    st, r, done = twin.classic_step("Acrobot-v1", [math.pi - 1e-3, 0.0, 0.0, 0.0], 1)
    assert done and r == 0.0
    # angles are wrapped into [-pi, pi], velocities bounded by 4 pi / 9 pi
    st, _, _ = twin.classic_step("Acrobot-v1", [3.1, 3.1, 12.0, 28.0], 2)
    assert -math.pi <= st[0] <= math.pi and -math.pi <= st[1] <= math.pi and abs(st[2]) <= 4 * math.pi and abs(st[3]) <= 9 * math.pi
    # torque-free swing conserves energy to the accuracy of RK4 over dt = 0.2
    def energy(s):
        t1, t2, w1, w2 = s
        y1 = -0.5 * math.cos(t1); y2 = -math.cos(t1) - 0.5 * math.cos(t1 + t2)
        v1sq = (0.5 * w1) ** 2
        vx2 = w1 * math.cos(t1) + 0.5 * (w1 + w2) * math.cos(t1 + t2); vy2 = w1 * math.sin(t1) + 0.5 * (w1 + w2) * math.sin(t1 + t2)
        return 0.5 * v1sq + 0.5 * (vx2 ** 2 + vy2 ** 2) + 0.5 * w1 ** 2 + 0.5 * (w1 + w2) ** 2 + 9.8 * (y1 + y2)
    s = np.array([0.4, -0.3, 0.0, 0.0]); e0 = energy(s)
    for _ in range(50):
        s, _, d = twin.classic_step("Acrobot-v1", s, 1)
    assert abs(energy(s) - e0) < 2e-3 * abs(e0)


@pytest.mark.parametrize("env", ENVS)
def test_twin_matches_python_restatement_step_by_step(twin, env):
    """Same actions, libm vs contract trigonometry: states agree to rounding noise (MountainCar) / to the chaotic
    amplification of 1-ulp differences over 200 RK4 steps (Acrobot)."""
    from oracle import pyref
    rng = np.random.default_rng(3)
    physics = pyref.mountaincar_physics if env == "MountainCar-v0" else pyref.acrobot_physics
    tol = 1e-13 if env == "MountainCar-v0" else 1e-9
    for trial in range(8):
        s_py = (float(rng.uniform(-0.6, -0.4)), 0.0) if env == "MountainCar-v0" else tuple(rng.uniform(-0.1, 0.1, 4))
        s_tw = np.array(s_py)
        for t in range(200):
            a = int(rng.integers(0, 3))
            s_py, r_py, d_py = physics(s_py, a)
            s_tw, r_tw, d_tw = twin.classic_step(env, s_tw, a)
            assert np.abs(np.array(s_py) - s_tw).max() <= tol and r_py == r_tw and d_py == d_tw
            if d_py:
                break


@pytest.mark.parametrize("env", ENVS)
def test_classic_rollout_golden(twin, golden, env):
    """Reference RolloutWorker + GymEnvModel over the Python restatement vs the twin: same returns (rtol 1e-4 asked by the
    north_star; equal here), same env-step counts, same actions and states (<= 1e-9) over the first 200 steps."""
    g = golden(GOLD[env])
    E, W, init = int(g["E"]), g["W"], g["init"]
    fit, steps = twin.population_classic(env, np.zeros((1, W.shape[1]), np.float32), n=W.shape[0], E=E, W_override=W, init=init)
    np.testing.assert_allclose(fit, g["fitness"], rtol=1e-4)
    assert np.mean(fit == g["fitness"]) >= 0.97 and np.mean(steps == g["steps"]) >= 0.97
    for j, i in enumerate(g["trace_ids"]):
        f, n, tr, ac = twin.rollout_classic(env, W[i], E=E, init=init, trace_steps=200)
        L = int(np.sum(g["trace_actions"][j] >= 0))
        assert np.array_equal(ac[:L], g["trace_actions"][j][:L])
        assert np.abs(tr[:L] - g["traces"][j][:L]).max() <= 1e-9


@pytest.mark.parametrize("env", ENVS)
def test_classic_init_streams(twin, env):
    a = twin.classic_init(env, 5, 0, 0, 0, 1)
    assert np.array_equal(a, twin.classic_init(env, 5, 0, 9, 4, 1))            # shared table ignores (gen, id)
    assert not np.array_equal(twin.classic_init(env, 5, 1, 0, 1, 1), twin.classic_init(env, 5, 1, 0, 2, 1))
    if env == "MountainCar-v0":
        assert -0.6 <= a[0] <= -0.4 and a[1] == 0.0
    else:
        assert np.all(np.abs(a) <= 0.1)


# ------------------------------------------------------------------------------------- Pendulum-v0, continuous-action head
def test_pendulum_known_answers(twin):
    """gym pendulum.py (Pendulum-v0): hanging down (th = pi) at rest with zero torque stays put and costs pi^2; upright at rest is
    an (unstable) equilibrium with zero cost; torque enters as 3 u dt; the speed is clipped at 8 after the angle has advanced."""
    st, r, done = twin.classic_step("Pendulum-v0", [math.pi, 0.0], 0.0)
    assert abs(st[0] - math.pi) < 1e-15 and abs(st[1]) < 1e-14 and abs(r + math.pi ** 2) < 1e-14 and not done
    st, r, done = twin.classic_step("Pendulum-v0", [0.0, 0.0], 0.0)
    assert abs(st[0]) < 1e-16 and abs(st[1]) < 1e-15 and r == 0.0 and not done
    st, r, _ = twin.classic_step("Pendulum-v0", [0.0, 0.0], 1.0)
    assert abs(st[1] - 3.0 * 0.05) < 1e-15 and abs(st[0] - 3.0 * 0.05 * 0.05) < 1e-16 and abs(r + 0.001) < 1e-18
    st, _, _ = twin.classic_step("Pendulum-v0", [math.pi / 2, 7.9], 1.0)       # gravity + torque push the speed past 8
    assert st[1] == 8.0 and st[0] > math.pi / 2 + 8.0 * 0.05                   # newth used the unclipped speed (v0 ordering)
    # angle_normalize: the cost sees the angle modulo 2 pi
    _, r1, _ = twin.classic_step("Pendulum-v0", [0.3, 0.0], 0.0)
    _, r2, _ = twin.classic_step("Pendulum-v0", [0.3 + 4 * math.pi, 0.0], 0.0)
    assert abs(r1 - r2) < 1e-14
    # the all-zero policy answers tanh(0) = 0: a free pendulum, 200 steps per episode
    f, n = twin.rollout_classic("Pendulum-v0", np.zeros(161, np.float32), E=2)
    assert n == 400 and f < 0.0


def test_pendulum_twin_matches_python_restatement(twin):
    from oracle import pyref
    rng = np.random.default_rng(4)
    for trial in range(8):
        s_py = (float(rng.uniform(-math.pi, math.pi)), float(rng.uniform(-1, 1)))
        s_tw = np.array(s_py)
        for t in range(200):
            u = float(np.float32(rng.uniform(-1, 1)))
            s_py, r_py, d_py = pyref.pendulum_physics(s_py, u)
            s_tw, r_tw, d_tw = twin.classic_step("Pendulum-v0", s_tw, u)
            assert np.abs(np.array(s_py) - s_tw).max() <= 1e-11 and abs(r_py - r_tw) <= 1e-11 and not d_py and not d_tw


def test_pendulum_rollout_golden(twin, golden):
    """The reference's RolloutWorker + GymEnvModel(num_state 3, num_action 1, discrete_action=False) -- the tanh head of
    networks/neural_network.py:32-33 -- over the Python Pendulum restatement vs the twin: returns within rtol 1e-4 (north_star),
    actions within float32 rounding noise of torch's tanh, states <= 1e-6 over the 200 steps of an episode."""
    g = golden("rollout_pendulum")
    E, W, init = int(g["E"]), g["W"], g["init"]
    fit, steps = twin.population_classic("Pendulum-v0", np.zeros((1, W.shape[1]), np.float32), n=W.shape[0], E=E, W_override=W, init=init)
    np.testing.assert_allclose(fit, g["fitness"], rtol=1e-4)
    assert np.array_equal(steps, g["steps"]) and np.all(steps == 200 * E)
    for j, i in enumerate(g["trace_ids"]):
        f, n, tr, ac = twin.rollout_classic("Pendulum-v0", W[i], E=E, init=init, trace_steps=200)
        assert ac.dtype == np.float32 and np.abs(ac - g["trace_actions"][j]).max() <= 2e-5
        assert np.abs(tr - g["traces"][j]).max() <= 1e-4


def test_pendulum_init_stream(twin):
    a = twin.classic_init("Pendulum-v0", 5, 0, 0, 0, 1)
    assert np.array_equal(a, twin.classic_init("Pendulum-v0", 5, 0, 9, 4, 1))
    assert abs(a[0]) <= math.pi and abs(a[1]) <= 1.0
    many = np.array([twin.classic_init("Pendulum-v0", 1, 1, 0, i, 0) for i in range(400)])
    assert many[:, 0].min() < -2.5 and many[:, 0].max() > 2.5 and many[:, 1].min() < -0.8 and many[:, 1].max() > 0.8
