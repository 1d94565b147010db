"""CartPole physics / wrapper semantics of the bit-twin vs the float64 Python restatement and
the reference-generated golden vectors (reference RolloutWorker + GymEnvModel over the shim)."""
import numpy as np

from oracle import pyref


def test_one_step_analytic(twin):
    # from rest, push right: closed form of Appendix A.1 with sin=0, cos=1
    st, done = twin.cartpole_step([0.0, 0.0, 0.0, 0.0], 1)
    temp = 10.0 / 1.1
    thacc = (-temp) / (0.5 * (4.0 / 3.0 - 0.1 / 1.1))
    xacc = temp - 0.05 * thacc / 1.1
    assert not done
    assert st[0] == 0.0 and st[2] == 0.0
    assert st[1] == 0.02 * xacc and st[3] == 0.02 * thacc
    st2, _ = twin.cartpole_step([0.0, 0.0, 0.0, 0.0], 0)
    assert st2[1] == -st[1] and st2[3] == -st[3]


def test_accelerations_equal_the_lagrangian_equations_of_motion(twin):
    """An independent pin of the physics (gym is not installable here): the restated closed forms for
    thetaacc / xacc must be THE solution of the cart-pole equations of motion derived from the Lagrangian of the system gym
    describes (cart 1.0 kg, uniform pole 0.1 kg of half-length 0.5 m pivoting on the cart, theta = 0 upright, no friction;
    Barto, Sutton & Anderson 1983):
        (M + m) x'' + m l cos(th) th''            = F + m l th'^2 sin(th)
        m l cos(th) x'' + (I + m l^2) th''        = m g l sin(th),      I = m (2 l)^2 / 12
    solved here as a 2x2 linear system per state; the twin's accelerations are read off one Euler step."""
    M, m, l, g, tau = 1.0, 0.1, 0.5, 9.8, 0.02
    inertia = m * (2 * l) ** 2 / 12.0
    rng = np.random.default_rng(11)
    for _ in range(2000):
        x, xd, th, thd = rng.uniform(-1, 1, 4) * np.array([2.4, 3.0, 0.2095, 3.5])
        a = int(rng.integers(0, 2))
        F = 10.0 if a == 1 else -10.0
        A = np.array([[M + m, m * l * np.cos(th)], [m * l * np.cos(th), inertia + m * l * l]])
        b = np.array([F + m * l * thd * thd * np.sin(th), m * g * l * np.sin(th)])
        xacc, thacc = np.linalg.solve(A, b)
        st, _ = twin.cartpole_step([x, xd, th, thd], a)
        assert abs((st[1] - xd) / tau - xacc) <= 1e-9 * max(1.0, abs(xacc))
        assert abs((st[3] - thd) / tau - thacc) <= 1e-9 * max(1.0, abs(thacc))
        assert st[0] == x + tau * xd and st[2] == th + tau * thd          # explicit Euler with the OLD velocities


def test_termination_thresholds(twin):
    assert twin.cartpole_step([2.39, 1.0, 0.0, 0.0], 1)[1] is True      # x crosses 2.4
    assert twin.cartpole_step([-2.39, -1.0, 0.0, 0.0], 0)[1] is True
    assert twin.cartpole_step([0.0, 0.0, 0.2090, 1.0], 1)[1] is True    # theta crosses 12 deg
    assert twin.cartpole_step([0.0, 0.0, 0.1, 0.0], 1)[1] is False


def test_single_steps_match_python_floats(twin):
    rng = np.random.default_rng(3)
    worst = 0.0
    for _ in range(20000):
        st = rng.uniform(-1, 1, 4) * np.array([2.4, 3.0, 0.2095, 3.5])
        a = int(rng.integers(0, 2))
        got, gd = twin.cartpole_step(st, a)
        want, wd = pyref.cartpole_physics(tuple(st), a)
        assert gd == wd
        worst = max(worst, np.abs(got - np.array(want)).max())
    assert worst < 2e-15


def test_trajectories_match_python_floats_200_steps(twin):
    """Same action sequence, |delta state| <= 1e-9 over the first 200 steps (north_star)."""
    rng = np.random.default_rng(4)
    for trial in range(40):
        st_t = rng.uniform(-0.05, 0.05, 4)
        st_p = tuple(st_t)
        a = 0
        for t in range(200):
            # bang-bang on the pole angle keeps the episode alive, with some random flips
            a = int(st_p[2] + 0.5 * st_p[3] > 0) if rng.random() > 0.05 else int(rng.integers(0, 2))
            st_t, d1 = twin.cartpole_step(st_t, a)
            st_p, d2 = pyref.cartpole_physics(st_p, a)
            assert np.abs(st_t - np.array(st_p)).max() <= 1e-9
            assert d1 == d2
            if d1:
                break


def _check_rollout_golden(twin, g):
    gru, pomdp, E = bool(g["gru"]), bool(g["pomdp"]), int(g["E"])
    W = g["W"]
    fit, steps = twin.population_cartpole(np.zeros((1, W.shape[1]), np.float32), gru=gru, pomdp=pomdp, n=W.shape[0],
                                          E=E, max_step=int(g["max_step"]), W_override=W, init=g["init"], nthreads=2)
    assert np.array_equal(steps / E, fit)
    # north_star: returns must match; near-tie action flips may touch <= 0.1 % of offspring
    assert (fit == g["fitness"]).mean() >= 0.999
    for j, i in enumerate(g["trace_ids"]):
        ref = g["traces"][j]
        n = int(np.isfinite(ref[:, 0]).sum())
        _, tr, ac = twin.rollout_cartpole(W[i], gru=gru, pomdp=pomdp, E=E, init=g["init"], trace_steps=200)
        assert np.array_equal(ac[:n], g["trace_actions"][j, :n])
        assert np.abs(tr[:n] - ref[:n]).max() <= 1e-9


def test_rollout_golden_mlp(twin, golden):
    _check_rollout_golden(twin, golden("rollout_cartpole_mlp"))


def large_golden(golden):
    """(init, W, reference returns) of tests/golden/rollout_cartpole_mlp_4096.npz; the weights are rebuilt from the seed exactly as
    oracle/make_golden.py::large_population drew them (numpy's legacy RandomState stream is frozen) and checked by CRC."""
    import zlib
    g = golden("rollout_cartpole_mlp_4096")
    P, seed = int(g["P"]), int(g["seed"])
    rng = np.random.RandomState(seed)
    init = rng.uniform(-0.05, 0.05, size=(5, 4))
    W = rng.normal(0, 2.0, size=(P, 226)).astype(np.float32)
    base = np.zeros(226, np.float32)
    base[:128].reshape(32, 4)[0] = [0.0, 0.5, 10.0, 3.0]
    base[160:224].reshape(2, 32)[1, 0] = 5.0
    base[160:224].reshape(2, 32)[0, 0] = -5.0
    W[P // 2:] = base + (W[P // 2:] * np.float32(0.15))
    assert zlib.crc32(np.ascontiguousarray(W).tobytes()) == int(g["w_crc32"]) and np.array_equal(init, g["init"])
    return init, W, g["fitness"]


def test_rollout_golden_mlp_4096_offspring(twin, golden):
    """the 99.9 % criterion on a population where it has teeth -- 4096 offspring, half random policies (short
    ragged episodes), half perturbations of a balancing parent (episodes up to the 500-step limit), returns of the reference's
    RolloutWorker + GymEnvModel (torch).  At most 4 offspring may differ (a near-tie action flip against torch's own tanh)."""
    init, W, want = large_golden(golden)
    fit, steps = twin.population_cartpole(np.zeros((1, 226), np.float32), n=W.shape[0], E=5, W_override=W, init=init, nthreads=8)
    same = fit == want
    assert same.mean() >= 0.999, (int((~same).sum()), np.flatnonzero(~same)[:10])
    assert np.array_equal(steps, (fit * 5).round().astype(np.int64))
    assert (want == 500).sum() > 1000 and (want < 30).sum() > 1000                 # both regimes are in the set


def test_rollout_golden_gru_pomdp(twin, golden):
    _check_rollout_golden(twin, golden("rollout_cartpole_gru_pomdp"))


def test_zero_policy_takes_action_zero(twin):
    """All-zero weights -> logits [0,0] -> argmax(softmax) = index 0 (neural_network.py:30-31)."""
    a, z, _ = twin.policy_step(np.zeros(226, np.float32), 4, 2, False, np.array([0.01, 0.0, 0.02, 0.0], np.float32))
    assert a == 0 and np.all(z == 0)
    t, tr, ac = twin.rollout_cartpole(np.zeros(226, np.float32), E=1, init=[[0.0, 0.0, 0.0, 0.0]], trace_steps=20)
    assert np.all(ac[:t] == 0) and 8 <= t <= 12


def test_max_step_truncation_and_pomdp(twin, golden):
    g = golden("rollout_cartpole_mlp")
    i = int(g["trace_ids"][0])
    t_full, _, _ = twin.rollout_cartpole(g["W"][i], E=1, init=g["init"], max_step=500)
    t_cut, _, _ = twin.rollout_cartpole(g["W"][i], E=1, init=g["init"], max_step=50)
    assert t_full > 50 and t_cut == 50
    # the POMDP mask changes what the policy sees (obs[1], obs[3] zeroed), so returns generally differ
    fit_a, _ = twin.population_cartpole(np.zeros((1, 226), np.float32), n=64, E=2, W_override=g["W"][:64], init=g["init"])
    fit_b, _ = twin.population_cartpole(np.zeros((1, 226), np.float32), pomdp=True, n=64, E=2, W_override=g["W"][:64],
                                        init=g["init"])
    assert not np.array_equal(fit_a, fit_b)


def test_philox_initial_states(twin):
    s0 = twin.cartpole_init(5, 0, 0, 0, 0)
    assert np.all(np.abs(s0) <= 0.05)
    # init_mode 0: shared by every offspring / generation (pool semantics); differs by episode
    assert np.array_equal(s0, twin.cartpole_init(5, 0, 9, 123, 0))
    assert not np.array_equal(s0, twin.cartpole_init(5, 0, 0, 0, 1))
    # init_mode 1: fresh per (generation, offspring, episode)
    assert not np.array_equal(twin.cartpole_init(5, 1, 0, 1, 0), twin.cartpole_init(5, 1, 0, 2, 0))
    assert not np.array_equal(twin.cartpole_init(5, 1, 0, 1, 0), twin.cartpole_init(5, 1, 1, 1, 0))
    allv = np.stack([twin.cartpole_init(5, 1, 0, i, 0) for i in range(4000)])
    assert abs(allv.mean()) < 2e-3 and abs(allv.std() - 0.1 / np.sqrt(12)) < 1e-3
