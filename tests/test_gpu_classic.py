"""GPU parity for the classic-control environments beyond CartPole-v1 (MountainCar-v0, Acrobot-v1, CartPole-v0):
the slot kernel through the C ABI against the CPU bit-twin (bit-exact under the numerical contract) and against the
reference-driven golden vectors (tolerances of the north_star written at each assert)."""
import os

import numpy as np
import pytest
import yaml

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SPEC = {"MountainCar-v0": (2, 3, 195, "rollout_mountaincar"), "Acrobot-v1": (6, 3, 323, "rollout_acrobot")}


def _engine(env, **kw):
    from simple_es_b200.engine import RolloutEngine
    obs, act, _, _ = SPEC[env]
    args = dict(env_name=env, obs_dim=obs, act_dim=act, gru=False, pomdp=False, max_step=None, eval_ep_num=5,
                population=1024, group=1024, n_head=1, n_parents=1, seed=0, init_mode="shared")
    args.update(kw)
    return RolloutEngine(**args)


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_sincos_full_bit_exact(twin):
    eng = _engine("Acrobot-v1", test_build=True)
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-100, 100, 2_000_000), rng.uniform(-4, 4, 1_000_000), np.arange(-80, 81) * (np.pi / 4), [0.0, -0.0]])
    s, c = twin.sincos_full(x)
    assert np.array_equal(eng.test_math("sin64_full", _cuda(x)).cpu().numpy(), s)
    assert np.array_equal(eng.test_math("cos64_full", _cuda(x)).cpu().numpy(), c)


@pytest.mark.parametrize("env", list(SPEC))
@pytest.mark.parametrize("E,init_mode,sigma", [(5, "shared", 2.0), (3, "fresh", 1.0), (1, "fresh", 3.0)])
def test_rollout_classic_philox_bit_exact(twin, env, E, init_mode, sigma):
    P = 1200
    D = SPEC[env][2]
    eng = _engine(env, population=P, group=P, eval_ep_num=E, seed=23, init_mode=init_mode)
    assert eng.D == D
    rng = np.random.default_rng(4)
    mu = rng.normal(0, 0.5, (1, D)).astype(np.float32)
    fit, steps = eng.rollout(3, sigma, _cuda(mu))
    tf, ts = twin.population_classic(env, mu, sigma=sigma, seed=23, gen=3, group=P, n_head=1, n=P, E=E,
                                     init_mode=0 if init_mode == "shared" else 1)
    assert np.array_equal(steps.cpu().numpy(), ts)
    assert np.array_equal(fit.cpu().numpy(), tf)            # float64 returns under the contract: bit-exact
    assert ts.min() >= E and len(np.unique(ts)) > 1 if env == "Acrobot-v1" else True


@pytest.mark.parametrize("env", list(SPEC))
def test_rollout_classic_verification_mode_matches_reference(twin, golden, env):
    """Golden = reference RolloutWorker + GymEnvModel over the float64 Python restatement (libm sin / cos).  The engine
    consumes the reference's weight arrays and initial states: returns within rtol 1e-4 (north_star) for every offspring
    whose action sequence is unchanged by the <= 1 ulp trigonometry difference (>= 97 % here), states of the traced
    episodes within 1e-9 over the first 200 steps, and bit-exact against the twin."""
    g = golden(SPEC[env][3])
    W, init, E = g["W"], g["init"], int(g["E"])
    P = W.shape[0]
    eng = _engine(env, population=P, group=P, eval_ep_num=E)
    fit, steps, trace, actions = eng.rollout(0, 0.0, None, w_override=_cuda(W), init_states=_cuda(init), n_trace=P)
    fit = fit.cpu().numpy(); trace = trace.cpu().numpy(); actions = actions.cpu().numpy()[:, :, 0]
    tf, ts = twin.population_classic(env, np.zeros((1, W.shape[1]), np.float32), n=P, E=E, W_override=W, init=init)
    assert np.array_equal(fit, tf) and np.array_equal(steps.cpu().numpy(), ts)
    np.testing.assert_allclose(fit, g["fitness"], rtol=1e-4)
    assert np.mean(fit == g["fitness"]) >= 0.97
    for j, i in enumerate(g["trace_ids"]):
        L = int(np.sum(g["trace_actions"][j] >= 0))
        assert np.array_equal(actions[i, :L], g["trace_actions"][j][:L])
        assert np.abs(trace[i, :L] - g["traces"][j][:L]).max() <= 1e-9
        f, n, tr, ac = twin.rollout_classic(env, W[i], E=E, init=init, trace_steps=200)
        assert np.array_equal(trace[i, :L], tr[:L])                 # vs the twin: bit-exact states


@pytest.mark.parametrize("n,E,max_step", [(1200, 5, 200), (65536, 3, 500), (4097, 32, 500)])
def test_rank_desc_integer_key_path_with_negative_fitness(twin, n, E, max_step):
    """K2's integer-key fast path (2 radix passes instead of 8) on the negative step totals of MountainCar / Acrobot:
    same permutation as the full float64 sort and as the oracle, ties by descending index."""
    rng = np.random.default_rng(n)
    totals = -rng.integers(0, E * max_step + 1, n)                   # heavy ties
    fit = totals.astype(np.float64) / E
    eng = _engine("Acrobot-v1", population=n, group=n, eval_ep_num=E, max_step=max_step)
    assert eng.key_bits == int(E * max_step).bit_length() and eng.key_scale == float(E)
    fast = eng.rank_desc(_cuda(fit)).cpu().numpy()
    full = eng.rank_desc(_cuda(fit), full_key=True).cpu().numpy()
    want = twin.rank_desc(fit)
    assert np.array_equal(fast, want) and np.array_equal(full, want)


def test_cartpole_v0_is_cartpole_with_a_200_step_limit(twin):
    from simple_es_b200.engine import RolloutEngine
    P, E, D = 512, 5, 226
    mu = np.zeros((1, D), np.float32)
    mu[0, :4] = [0.0, 0.5, 10.0, 3.0]; mu[0, 160 + 32] = 5.0; mu[0, 160] = -5.0      # a balancing parent
    eng = RolloutEngine("CartPole-v0", 4, 2, False, False, 500, E, P, P, 1, 1, seed=2)
    assert eng.max_step == 200
    fit, steps = eng.rollout(0, 0.05, _cuda(mu))
    tf, ts = twin.population_cartpole(mu, sigma=0.05, seed=2, gen=0, group=P, n_head=1, n=P, E=E, max_step=200, nthreads=8)
    assert np.array_equal(steps.cpu().numpy(), ts) and np.array_equal(fit.cpu().numpy(), tf)
    assert fit.max().item() == 200.0


def test_classic_engine_rejects_wrong_shapes():
    from simple_es_b200.engine import RolloutEngine
    with pytest.raises(ValueError, match="num_state=2"):
        RolloutEngine("MountainCar-v0", 4, 2, False, False, 200, 5, 64, 64, 1, 1)
    with pytest.raises(RuntimeError, match="only CartPole supports pomdp"):          # envs/gym_wrapper.py:11-19
        RolloutEngine("Acrobot-v1", 6, 3, False, True, 500, 5, 64, 64, 1, 1)


@pytest.mark.parametrize("conf,env,ngen", [("mountaincar.yaml", "MountainCar-v0", 3), ("acrobot.yaml", "Acrobot-v1", 3)])
def test_classic_loops_match_oracle_composition(twin, conf, env, ngen):
    """conf/mountaincar.yaml (simple_genetic) and conf/acrobot.yaml (openai_es) through B200Loop: every generation's
    fitness vector, rank order and updated parameters equal a composition of oracle steps (negative float64 fitness goes
    through K2's full-key radix path)."""
    from simple_es_b200.loop import B200Loop
    cfg = yaml.load(open(os.path.join(ROOT, "conf", conf)), Loader=yaml.FullLoader)
    D = SPEC[env][2]
    E = 3
    if cfg["strategy"]["name"] == "simple_genetic":
        cfg["strategy"].update(offspring_num=400, elite_num=8, init_sigma=2.0, sigma_decay=0.5)
        loop = B200Loop(cfg, ngen, 1, E, save_model_period=0, seed=5, quiet=True)
        s = loop.strategy
        k, grp = 8, 50
        P = k * grp
        elites = np.zeros((k, D), np.float32)
        sig_pop, sig_rep = 2.0, 2.0
        for gen in range(ngen):
            s.step()
            tf, ts = twin.population_classic(env, elites, sigma=sig_pop, seed=5, gen=gen, group=grp, n_head=1, n=P, E=E, init_mode=1)
            assert np.array_equal(s.fitness.cpu().numpy(), tf)
            order = twin.rank_desc(tf)
            assert np.array_equal(s.order.cpu().numpy(), order)
            elites = twin.materialize(elites, sig_pop, 5, gen, grp, 1, order[:k])
            assert np.array_equal(s.parents.cpu().numpy(), elites)
            sig_pop = sig_rep
            sig_rep *= 0.5
    else:
        cfg["strategy"].update(offspring_num=512, init_sigma=0.5, learning_rate=0.1, sigma_decay=0.9)
        loop = B200Loop(cfg, ngen, 1, E, save_model_period=0, seed=5, quiet=True)
        s = loop.strategy
        P, sigma, lr = 512, 0.5, 0.1
        mu = np.zeros((1, D), np.float32); m = np.zeros(D, np.float32); v = np.zeros(D, np.float32)
        for gen in range(ngen):
            s.step()
            tf, ts = twin.population_classic(env, mu, sigma=sigma, seed=5, gen=gen, group=P, n_head=1, n=P, E=E, init_mode=0)
            assert np.array_equal(s.fitness.cpu().numpy(), tf)
            order = twin.rank_desc(tf)
            assert np.array_equal(s.order.cpu().numpy(), order)
            shaped = twin.centered_rank(order)
            g = twin.grad_openai(shaped, D, 5, gen, P, 1, -(lr / (P * sigma)))
            mu1, m, v = twin.adam(mu[0], m, v, g, s.engine.adam_a(lr, gen + 1))
            mu = mu1[None]
            assert np.array_equal(s.parents.cpu().numpy(), mu)
            sigma *= 0.9


# ------------------------------------------------------------------------------------- Pendulum-v0, continuous-action head
def _pendulum(**kw):
    from simple_es_b200.engine import RolloutEngine
    args = dict(env_name="Pendulum-v0", obs_dim=3, act_dim=1, gru=False, pomdp=False, max_step=200, eval_ep_num=5, population=1024,
                group=1024, n_head=1, n_parents=1, seed=0, init_mode="shared", discrete_action=False)
    args.update(kw)
    return RolloutEngine(**args)


@pytest.mark.parametrize("E,init_mode,sigma", [(5, "shared", 1.0), (3, "fresh", 0.5), (1, "fresh", 2.0)])
def test_rollout_pendulum_philox_bit_exact(twin, E, init_mode, sigma):
    """`discrete_action: False` (the tanh head, networks/neural_network.py:32-33) + Pendulum-v0 physics in the slot kernel:
    float64 returns bit-exact against the twin."""
    P = 1500
    eng = _pendulum(population=P, group=P, eval_ep_num=E, seed=29, init_mode=init_mode)
    assert eng.D == 161 and eng.key_bits == 0
    mu = np.random.default_rng(6).normal(0, 0.5, (1, 161)).astype(np.float32)
    fit, steps = eng.rollout(4, sigma, _cuda(mu))
    tf, ts = twin.population_classic("Pendulum-v0", mu, sigma=sigma, seed=29, gen=4, group=P, n_head=1, n=P, E=E,
                                     init_mode=0 if init_mode == "shared" else 1)
    assert np.array_equal(steps.cpu().numpy(), ts) and np.all(ts == 200 * E)
    assert np.array_equal(fit.cpu().numpy(), tf)


def test_rollout_pendulum_verification_mode_matches_reference(twin, golden):
    """Golden = the reference's RolloutWorker + GymEnvModel(3, 1, discrete_action=False) over the float64 Python Pendulum
    restatement.  The engine consumes the reference's weights and initial states: returns within rtol 1e-4 (north_star),
    traced actions within float32 rounding noise of torch's tanh, traced states within 1e-4 over the whole 200-step episode;
    bit-exact against the twin."""
    g = golden("rollout_pendulum")
    W, init, E = g["W"], g["init"], int(g["E"])
    P = W.shape[0]
    eng = _pendulum(population=P, group=P, eval_ep_num=E)
    fit, steps, trace, actions = eng.rollout(0, 0.0, None, w_override=_cuda(W), init_states=_cuda(init), n_trace=P)
    fit = fit.cpu().numpy(); trace = trace.cpu().numpy(); actions = actions.cpu().numpy()[:, :, 0].copy().view(np.float32)
    tf, ts = twin.population_classic("Pendulum-v0", np.zeros((1, 161), np.float32), n=P, E=E, W_override=W, init=init)
    assert np.array_equal(fit, tf) and np.array_equal(steps.cpu().numpy(), ts)
    np.testing.assert_allclose(fit, g["fitness"], rtol=1e-4)
    for j, i in enumerate(g["trace_ids"]):
        assert np.abs(actions[i] - g["trace_actions"][j]).max() <= 2e-5
        assert np.abs(trace[i] - g["traces"][j]).max() <= 1e-4
        f, n, tr, ac = twin.rollout_classic("Pendulum-v0", W[i], E=E, init=init, trace_steps=200)
        assert np.array_equal(trace[i], tr) and np.array_equal(actions[i], ac)


def test_pendulum_loop_matches_oracle_composition(twin):
    """conf/pendulum.yaml (openai_es, discrete_action: False) through B200Loop: fitness, rank order (float64 keys, full radix
    path) and updated parameters of every generation equal a composition of oracle steps; the checkpoint has the reference's
    state_dict keys with a 1-row fc2."""
    from simple_es_b200.loop import B200Loop
    cfg = yaml.load(open(os.path.join(ROOT, "conf", "pendulum.yaml")), Loader=yaml.FullLoader)
    cfg["strategy"].update(offspring_num=512, init_sigma=0.5, learning_rate=0.1, sigma_decay=0.9)
    D, E, P, sigma, lr = 161, 3, 512, 0.5, 0.1
    loop = B200Loop(cfg, 3, 1, E, save_model_period=0, seed=5, quiet=True)
    s = loop.strategy
    mu = np.zeros((1, D), np.float32); m = np.zeros(D, np.float32); v = np.zeros(D, np.float32)
    for gen in range(3):
        s.step()
        tf, ts = twin.population_classic("Pendulum-v0", mu, sigma=sigma, seed=5, gen=gen, group=P, n_head=1, n=P, E=E, init_mode=1)
        assert np.array_equal(s.fitness.cpu().numpy(), tf)
        order = twin.rank_desc(tf)
        assert np.array_equal(s.order.cpu().numpy(), order)
        g = twin.grad_openai(twin.centered_rank(order), D, 5, gen, P, 1, -(lr / (P * sigma)))
        mu1, m, v = twin.adam(mu[0], m, v, g, s.engine.adam_a(lr, gen + 1))
        mu = mu1[None]
        assert np.array_equal(s.parents.cpu().numpy(), mu)
        sigma *= 0.9
    sd = loop.elite_state_dict()
    assert list(sd) == ["fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"] and tuple(sd["fc2.weight"].shape) == (1, 32)


def test_continuous_head_is_rejected_elsewhere():
    from simple_es_b200.engine import RolloutEngine
    with pytest.raises(ValueError, match="discrete_action"):
        RolloutEngine("CartPole-v1", 4, 2, False, False, 500, 5, 64, 64, 1, 1, discrete_action=False)
    with pytest.raises(ValueError, match="discrete_action"):
        RolloutEngine("Pendulum-v0", 3, 1, False, False, 200, 5, 64, 64, 1, 1)
