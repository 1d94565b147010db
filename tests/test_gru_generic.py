"""The recurrent policy (`gru: True`) on every environment other than CartPole: MountainCar-v0, Acrobot-v1, Pendulum-v0
(continuous head) and simple_spread N = 2 / 3 with ONE HIDDEN STATE PER AGENT COPY, as the reference builds them
(learning_strategies/evolution/utils.py:4-8 wrap_agentid; networks/neural_network.py:15-16,25-27,38-40).

CPU part: the C twin against golden vectors produced by the reference's own RolloutWorker + GymEnvModel(gru=True) + wrap_agentid
(oracle/make_golden.py), and the generic kernel (csrc/rollout_gru_generic.cuh) on the host SIMT emulator against the twin, bit for
bit.  GPU part (-m gpu): the same kernel on a B200 through the C ABI."""
import numpy as np
import pytest

CASES = {
    # name: (env, obs, act, n_agents, golden)
    "mountaincar": ("MountainCar-v0", 2, 3, 1, "rollout_mountaincar_gru"),
    "acrobot": ("Acrobot-v1", 6, 3, 1, "rollout_acrobot_gru"),
    "pendulum": ("Pendulum-v0", 3, 1, 1, "rollout_pendulum_gru"),
    "spread2": ("simple_spread", 12, 5, 2, "rollout_spread_n2_gru"),
    "spread3": ("simple_spread", 18, 5, 3, "rollout_spread_n3_gru"),
}


def _twin_population(twin, env, N, mu, **kw):
    if env == "simple_spread":
        kw.pop("nthreads", None)
        return twin.population_mpe(mu, N=N, gru=True, **kw)
    return twin.population_classic(env, mu, gru=True, **kw)


def _twin_rollout(twin, env, N, w, E, init, steps):
    if env == "simple_spread":
        return twin.rollout_mpe(w, N=N, E=E, init=init, trace_steps=steps, gru=True)
    return twin.rollout_classic(env, w, E=E, init=init, trace_steps=steps, gru=True)


# ------------------------------------------------------------------------------------- twin vs the reference's classes
@pytest.mark.parametrize("case", list(CASES))
def test_twin_matches_reference_goldens(twin, golden, case):
    """Returns within rtol 1e-4 (north_star); discrete actions of the traced episodes equal, states within 1e-9 (Pendulum's
    continuous torque: float32 rounding noise of torch's tanh / sigmoid against the contract's, states within 1e-4)."""
    env, obs, act, N, name = CASES[case]
    g = golden(name)
    W, init, E = g["W"], g["init"], int(g["E"])
    fit, steps = _twin_population(twin, env, N, np.zeros((1, W.shape[1]), np.float32), n=W.shape[0], E=E, W_override=W, init=init)
    np.testing.assert_allclose(fit, g["fitness"], rtol=1e-4)
    T = 25 if env == "simple_spread" else 200
    ids = range(g["traces"].shape[0]) if env == "simple_spread" else g["trace_ids"]
    for j, i in enumerate(ids):
        f, n, tr, ac = _twin_rollout(twin, env, N, W[i], E, init, T)
        ga, gt = g["trace_actions"][j], g["traces"][j]
        if env == "Pendulum-v0":
            # a continuous torque fed back through a recurrent state compounds the float32 differences over the 200 steps
            assert np.abs(ac - ga).max() <= 5e-4 and np.abs(tr - gt).max() <= 5e-3
        else:
            L = T if env == "simple_spread" else int(np.sum(ga >= 0))
            assert np.array_equal(np.asarray(ac)[:L].reshape(L, -1), np.asarray(ga)[:L].reshape(L, -1))
            assert np.abs(tr[:L] - gt[:L]).max() <= 1e-9


def test_per_agent_hidden_state_matters(twin, golden):
    """wrap_agentid's deep copies give every agent its own recurrent state; the golden (the reference's separate copies) is what
    the twin's per-agent hidden states reproduce, and the recurrent run is not the feed-forward policy of the same fc1 / fc2."""
    g = golden("rollout_spread_n2_gru")
    W, init, E = g["W"], g["init"], int(g["E"])
    w = W[1]
    f, n, tr, ac = twin.rollout_mpe(w, N=2, E=1, init=init[:1], trace_steps=25, gru=True)
    assert np.array_equal(ac, g["trace_actions"][1])
    # the MLP-only part of the same weights is a different policy altogether; the GRU run must not degenerate to it
    D_mlp = 12 * 32 + 32 + 5 * 32 + 5
    w_mlp = np.concatenate([w[:12 * 32 + 32], w[-(5 * 32 + 5):]]).astype(np.float32)
    assert w_mlp.size == D_mlp
    _, _, _, ac_mlp = twin.rollout_mpe(w_mlp, N=2, E=1, init=init[:1], trace_steps=25)
    assert not np.array_equal(ac_mlp, ac)


# ------------------------------------------------------------------------------------- emulated kernel vs twin (CPU)
@pytest.fixture(scope="module")
def emu():
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    from simt_emu import emu_engine
    emu_engine.load()
    return emu_engine.EmuEngine


@pytest.mark.parametrize("case,P,E,init_mode,sigma", [("mountaincar", 20, 3, "fresh", 1.5), ("mountaincar", 10, 7, "shared", 1.0),
                                                      ("acrobot", 8, 2, "fresh", 1.0), ("pendulum", 16, 3, "fresh", 0.7),
                                                      ("pendulum", 8, 6, "shared", 0.5), ("spread2", 16, 3, "fresh", 0.7),
                                                      ("spread2", 8, 5, "shared", 0.5), ("spread3", 8, 2, "fresh", 0.7),
                                                      ("spread3", 6, 3, "shared", 0.5)])
def test_emu_gru_generic_philox_bit_exact(emu, twin, case, P, E, init_mode, sigma):
    env, obs, act, N, _ = CASES[case]
    eng = emu(env, obs, act, gru=True, population=P, group=P, n_head=1, eval_ep_num=E, seed=5, init_mode=init_mode, max_step=None, n_agents=N)
    mu = np.random.default_rng(1).normal(0, 0.3, (1, eng.D)).astype(np.float32)
    fit, steps = eng.rollout(3, sigma, mu)
    tf, ts = _twin_population(twin, env, N, mu, sigma=sigma, seed=5, gen=3, group=P, n_head=1, n=P, E=E,
                              init_mode=0 if init_mode == "shared" else 1, nthreads=4)
    assert np.array_equal(steps, ts) and np.array_equal(fit, tf)


@pytest.mark.parametrize("case", list(CASES))
def test_emu_gru_generic_traces_equal_the_twin(emu, twin, golden, case):
    env, obs, act, N, name = CASES[case]
    g = golden(name)
    n = min(6, g["W"].shape[0])
    W, init, E = g["W"][:n], g["init"], int(g["E"])
    eng = emu(env, obs, act, gru=True, population=n, group=n, eval_ep_num=E, max_step=None, n_agents=N)
    fit, steps, trace, actions = eng.rollout(0, 0.0, None, w_override=W, init_states=init, n_trace=n)
    T = 25 if env == "simple_spread" else 200
    for j in range(n):
        tf, tsteps, ttr, tac = _twin_rollout(twin, env, N, W[j], E, init, T)
        m = int(np.isfinite(ttr[:, 0]).sum())
        assert m > 0 and np.array_equal(trace[j, :m], ttr[:m])
        got = actions[j, :m]
        if env == "Pendulum-v0":
            assert np.array_equal(got[:, 0].copy().view(np.float32), tac[:m])
        else:
            assert np.array_equal(got.reshape(m, -1), np.asarray(tac)[:m].reshape(m, -1))
        assert fit[j] == tf and steps[j] == tsteps
    np.testing.assert_allclose(fit, g["fitness"][:n], rtol=1e-4)             # vs the reference path


# ------------------------------------------------------------------------------------- B200 (-m gpu)
def _cuda(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("case,P,E,init_mode,sigma", [("mountaincar", 600, 3, "fresh", 1.5), ("acrobot", 300, 2, "fresh", 1.0),
                                                      ("pendulum", 600, 5, "shared", 0.7), ("spread2", 700, 5, "fresh", 0.7),
                                                      ("spread3", 500, 3, "fresh", 0.7), ("spread2", 300, 7, "shared", 0.5)])
def test_gpu_gru_generic_philox_bit_exact(twin, case, P, E, init_mode, sigma):
    torch = pytest.importorskip("torch")
    from simple_es_b200.engine import RolloutEngine
    env, obs, act, N, _ = CASES[case]
    eng = RolloutEngine(env, obs, act, True, False, None, E, P, P, 1, 1, seed=5, init_mode=init_mode, n_agents=N,
                        discrete_action=env != "Pendulum-v0")
    mu = np.random.default_rng(1).normal(0, 0.3, (1, eng.D)).astype(np.float32)
    fit, steps = eng.rollout(3, sigma, _cuda(mu))
    tf, ts = _twin_population(twin, env, N, mu, sigma=sigma, seed=5, gen=3, group=P, n_head=1, n=P, E=E,
                              init_mode=0 if init_mode == "shared" else 1, nthreads=8)
    assert np.array_equal(steps.cpu().numpy(), ts) and np.array_equal(fit.cpu().numpy(), tf)


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(CASES))
def test_gpu_gru_generic_verification_mode_matches_reference(twin, golden, case):
    """The engine consumes the reference's weight arrays and initial states (goldens of the reference's RolloutWorker +
    GymEnvModel(gru=True) + wrap_agentid): returns within rtol 1e-4 (north_star); traces bit-exact against the twin."""
    torch = pytest.importorskip("torch")
    from simple_es_b200.engine import RolloutEngine
    env, obs, act, N, name = CASES[case]
    g = golden(name)
    W, init, E = g["W"], g["init"], int(g["E"])
    P = W.shape[0]
    eng = RolloutEngine(env, obs, act, True, False, None, E, P, P, 1, 1, n_agents=N, discrete_action=env != "Pendulum-v0")
    fit, steps, trace, actions = eng.rollout(0, 0.0, None, w_override=_cuda(W), init_states=_cuda(init), n_trace=P)
    fit = fit.cpu().numpy(); trace = trace.cpu().numpy(); actions = actions.cpu().numpy()
    tf, ts = _twin_population(twin, env, N, np.zeros((1, W.shape[1]), np.float32), n=P, E=E, W_override=W, init=init)
    assert np.array_equal(fit, tf) and np.array_equal(steps.cpu().numpy(), ts)
    np.testing.assert_allclose(fit, g["fitness"], rtol=1e-4)
    T = 25 if env == "simple_spread" else 200
    for j in range(min(P, 6)):
        f, n, ttr, tac = _twin_rollout(twin, env, N, W[j], E, init, T)
        m = int(np.isfinite(ttr[:, 0]).sum())
        assert np.array_equal(trace[j, :m], ttr[:m])
        if env == "Pendulum-v0":
            assert np.array_equal(actions[j, :m, 0].copy().view(np.float32), tac[:m])
        else:
            assert np.array_equal(actions[j, :m].reshape(m, -1), np.asarray(tac)[:m].reshape(m, -1))
