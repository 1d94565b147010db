"""GPU parity: the sm_100a kernels, called through the C ABI (simple_es_b200.engine -> ctypes ->
libses_b200.so), against the CPU bit-twin oracle on the same seeded inputs and against the
reference-generated golden fixtures.  Integer / index results and everything under the numerical
contract must be bit-exact; comparisons with the reference's torch/numpy path carry the
tolerances the north_star states (written next to each assert)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

D = 226


def _engine(**kw):
    from simple_es_b200.engine import RolloutEngine
    args = dict(env_name="CartPole-v1", obs_dim=4, act_dim=2, gru=False, pomdp=False, max_step=500, eval_ep_num=5,
                population=1024, group=1024, n_head=1, n_parents=1, seed=0, init_mode="shared")
    args.update(kw)
    return RolloutEngine(**args)


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ------------------------------------------------------------------------------------- contract
def test_math_contract_bit_exact(twin):
    eng = _engine(test_build=True)            # the contract functions are reachable through the test build's hooks only
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-10, 10, 1_000_000), rng.normal(0, 1, 1_000_000), [0.0, -0.0, 50, -50]]).astype(np.float32)
    assert np.array_equal(eng.test_math("tanh", _cuda(x)).cpu().numpy(), twin.tanhf(x))
    assert np.array_equal(eng.test_math("sigmoid", _cuda(x)).cpu().numpy(), twin.sigmf(x))
    # K1 runs tanh with the division fast path written out; it must be the same function, for EVERY float
    assert np.array_equal(eng.test_math("tanh_fast", _cuda(x)).cpu().numpy(), twin.tanhf(x))
    assert eng.test_tanh_fast_exhaustive(0.0, 3.0e38) == 0
    # the packed FFMA2 forms K1 runs on pairs of hidden units (variant 2, the default, has no Newton step on the
    # reciprocal seed; variant 1 keeps it): the same function in both halves, for EVERY float
    assert eng.test_tanh_x2_exhaustive(False, 0.0, 3.0e38) == 0
    assert eng.test_tanh_x2_exhaustive(True, 0.0, 3.0e38) == 0
    # ... and its 3-instruction division by total_mass must be the IEEE quotient (2^33 random operands)
    assert eng.test_div_total_mass(1 << 33) == 0
    u = ((rng.integers(0, 2 ** 24, 1_000_000) + 0.5) * 2.0 ** -24).astype(np.float32)
    assert np.array_equal(eng.test_math("ln", _cuda(u)).cpu().numpy(), twin.lnf(u))
    v = (rng.integers(0, 2 ** 24, 1_000_000) * 2.0 ** -24).astype(np.float32)
    s, c = twin.sincos2pif(v)
    assert np.array_equal(eng.test_math("sin2pi", _cuda(v)).cpu().numpy(), s)
    assert np.array_equal(eng.test_math("cos2pi", _cuda(v)).cpu().numpy(), c)
    th = rng.uniform(-0.5, 0.5, 1_000_000)
    s, c = twin.sincos(th)
    assert np.array_equal(eng.test_math("sin64", _cuda(th)).cpu().numpy(), s)
    assert np.array_equal(eng.test_math("cos64", _cuda(th)).cpu().numpy(), c)


def test_philox_normals_bit_exact(twin):
    eng = _engine(seed=1234, test_build=True)
    for gen, idx in [(0, 0), (0, 1), (3, 77), (1000, 65535), (2 ** 31, 2 ** 20 - 1)]:
        got = eng.test_normals(gen, idx).cpu().numpy()
        assert np.array_equal(got, twin.normals(1234, gen, idx, D))


@pytest.mark.parametrize("strategy,n,k", [("simple_evolution", 96, 10), ("openai_es", 128, None), ("simple_genetic", 100, 8)])
def test_materialize_layouts_bit_exact(twin, strategy, n, k):
    from simple_es_b200.engine import population_layout
    P, group, n_head, n_par = population_layout(strategy, n, k)
    eng = _engine(population=P, group=group, n_head=n_head, n_parents=n_par, seed=5)
    rng = np.random.default_rng(1)
    parents = rng.normal(0, 1, (n_par, D)).astype(np.float32)
    ids = np.arange(P, dtype=np.int32)
    got = eng.materialize(7, 0.37, _cuda(parents), _cuda(ids)).cpu().numpy()
    want = twin.materialize(parents, 0.37, 5, 7, group, n_head, ids)
    assert np.array_equal(got, want)
    heads = [i for i in range(P) if i % group < n_head]
    assert all(np.array_equal(got[i], parents[i // group]) for i in heads)


# ------------------------------------------------------------------------------------- K1
@pytest.mark.parametrize("init_mode,pomdp,E", [("shared", False, 5), ("fresh", False, 3), ("shared", True, 5), ("shared", False, 1)])
def test_rollout_philox_bit_exact(twin, init_mode, pomdp, E):
    P = 3000
    eng = _engine(population=P, group=P, n_head=1, eval_ep_num=E, pomdp=pomdp, init_mode=init_mode, seed=11)
    mu = np.zeros((1, D), np.float32)
    fit, steps = eng.rollout(2, 2.0, _cuda(mu))
    tf, ts = twin.population_cartpole(mu, pomdp=pomdp, sigma=2.0, seed=11, gen=2, group=P, n_head=1, n=P, E=E,
                                      init_mode=0 if init_mode == "shared" else 1, nthreads=8)
    assert np.array_equal(steps.cpu().numpy(), ts)          # integer returns: bit-exact
    assert np.array_equal(fit.cpu().numpy(), tf)
    assert ts.max() > 3 * ts.min()                          # the case really has ragged episode lengths


@pytest.mark.parametrize("knobs,E", [({"SES_K1_VARIANT": "0"}, 5), ({"SES_K1_VARIANT": "1"}, 5), ({"SES_K1_VARIANT": "2"}, 5),
                                     ({"SES_K1_VARIANT": "3"}, 5), ({"SES_K1_VARIANT": "4"}, 5), ({"SES_K1_VARIANT": "5"}, 5),
                                     ({"SES_K1_VARIANT": "4"}, 3), ({"SES_K1_VARIANT": "4"}, 1)])
def test_rollout_k1_variants_bit_exact(twin, knobs, E, monkeypatch):
    """Every K1 code path is the same function: variant 0 (scalar FFMA, flat slot table), 1 / 2 (packed FFMA2 over
    hidden-unit pairs, permuted slot table, with / without the Newton step of the tanh division), 3 / 4 / 5 (variant 2
    with W2+b2 / W2+b2+b1 / b1 of the lane's slot held in registers; 4 is the default)
    must all reproduce the oracle bit for bit: Philox and verification (w_override) paths, ragged and 500-step
    episodes, 8 / 16 / 32 slots per warp (E = 5 / 3 / 1)."""
    monkeypatch.setenv("SES_B200_TEST_BUILD", "1")        # the alternative kernels live in libses_b200_tests.so
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    P = 2048
    rng = np.random.default_rng(5)
    for sigma, seed, pomdp in [(2.0, 21, False), (0.05, 22, False), (1.0, 23, True)]:
        mu = np.zeros((1, D), np.float32)
        if sigma < 1.0:
            w1 = mu[0, :128].reshape(32, 4); w2 = mu[0, 160:224].reshape(2, 32)
            w1[0] = [0.0, 0.5, 10.0, 3.0]; w2[1, 0] = 5.0; w2[0, 0] = -5.0
        eng = _engine(population=P, group=P, n_head=1, eval_ep_num=E, seed=seed, pomdp=pomdp)
        fit, steps = eng.rollout(1, sigma, _cuda(mu))
        tf, ts = twin.population_cartpole(mu, pomdp=pomdp, sigma=sigma, seed=seed, gen=1, group=P, n_head=1, n=P, E=E, nthreads=8)
        assert np.array_equal(steps.cpu().numpy(), ts) and np.array_equal(fit.cpu().numpy(), tf)
        eng.close()
    W = rng.normal(0, 1.5, (256, D)).astype(np.float32)
    init = rng.uniform(-0.05, 0.05, (E, 4))
    eng = _engine(population=256, group=256, n_head=1, eval_ep_num=E)
    fit, steps = eng.rollout(0, 0.0, None, w_override=_cuda(W), init_states=_cuda(init))
    tf, ts = twin.population_cartpole(np.zeros((1, D), np.float32), n=256, group=256, E=E, W_override=W, init=init, nthreads=8)
    assert np.array_equal(steps.cpu().numpy(), ts) and np.array_equal(fit.cpu().numpy(), tf)


def test_rollout_trained_parent_long_episodes(twin):
    """A parent that balances the pole (bang-bang on theta + theta_dot) with small sigma: most
    episodes hit the 500-step limit, exercising long trajectories and the truncation rule."""
    P, E = 1024, 5
    mu = np.zeros((1, D), np.float32)
    w1 = mu[0, :128].reshape(32, 4); w2 = mu[0, 160:224].reshape(2, 32)
    w1[0] = [0.0, 0.5, 10.0, 3.0]                           # h0 = tanh(.5 xdot + 10 theta + 3 thetadot)
    w2[1, 0] = 5.0; w2[0, 0] = -5.0
    eng = _engine(population=P, group=P, n_head=1, eval_ep_num=E, seed=3)
    fit, steps = eng.rollout(0, 0.05, _cuda(mu))
    tf, ts = twin.population_cartpole(mu, sigma=0.05, seed=3, gen=0, group=P, n_head=1, n=P, E=E, nthreads=8)
    assert np.array_equal(steps.cpu().numpy(), ts)
    assert (ts == 500 * E).mean() > 0.5


@pytest.mark.parametrize("knobs", [{}, {"SES_K1_TAIL": "0"}, {"SES_K1_TAIL": "1"}, {"SES_K1_TAIL": "100"}, {"SES_K1_SPLIT": "0"},
                                   {"SES_K1_SPARSE": "0"}, {"SES_ROLLOUT_LANES": "30"}, {"SES_ROLLOUT_LANES": "7"},
                                   {"SES_K1_SPARSE_RANK": "0", "SES_K1_SPARSE_QUOTA": "3"}, {"SES_ROLLOUT_CTAS_PER_SM": "1"}])
def test_rollout_launch_geometry_never_changes_a_bit(twin, knobs, monkeypatch):
    """Round 2 scheduler of the CartPole-MLP kernel: all 32 lanes with the exact-request queue for the last rounds (an offspring's
    episodes may run in two warps and are added up with integer atomics), the straggler phase (one episode on 2 / 4 lanes with
    mirrored speculative physics), sparse warps (a launch that does not fill the SMs) -- whatever the geometry, fitness and
    env-step counts are the twin's, for ragged and for 500-step populations, several population sizes around the round
    boundaries, twice in a row (the cross-warp accumulators must come back to zero)."""
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    w1 = np.zeros((1, D), np.float32)
    w1[0, :4] = [0.0, 0.5, 10.0, 3.0]; w1[0, 160 + 32] = 5.0; w1[0, 160] = -5.0
    for P, E, sigma, parent in [(8192, 5, 0.05, w1), (13000, 5, 0.05, w1), (3000, 5, 2.0, np.zeros((1, D), np.float32)),
                                (4097, 3, 0.05, w1), (2500, 7, 0.3, w1)]:
        eng = _engine(population=P, group=P, n_head=1, eval_ep_num=E, seed=31)
        tf = ts = None
        for gen in (0, 1):
            fit, steps = eng.rollout(gen, sigma, _cuda(parent))
            tf, ts = twin.population_cartpole(parent, sigma=sigma, seed=31, gen=gen, group=P, n_head=1, n=P, E=E, nthreads=8)
            assert np.array_equal(steps.cpu().numpy(), ts) and np.array_equal(fit.cpu().numpy(), tf)
        eng.close()


def test_rollout_4096_offspring_match_reference_returns(twin, golden):
    """The reference's own weight arrays at population 4096 (tests/golden/rollout_cartpole_mlp_4096.npz: returns of the
    reference's RolloutWorker + GymEnvModel): >= 99.9 % of the returns exactly equal (north_star), and every return equal to
    the twin's."""
    from test_oracle_cartpole import large_golden
    init, W, want = large_golden(golden)
    P = W.shape[0]
    eng = _engine(population=P, group=P, n_head=1, eval_ep_num=5)
    fit, steps = eng.rollout(0, 0.0, None, w_override=_cuda(W), init_states=_cuda(init))
    fit = fit.cpu().numpy()
    assert (fit == want).mean() >= 0.999
    tf, ts = twin.population_cartpole(np.zeros((1, 226), np.float32), n=P, E=5, W_override=W, init=init, nthreads=8)
    assert np.array_equal(fit, tf) and np.array_equal(steps.cpu().numpy(), ts)


def test_rollout_verification_mode_matches_reference(twin, golden):
    """The engine consumes the reference's own weight arrays and initial states (north_star
    verification mode).  Golden = reference RolloutWorker + GymEnvModel (torch CPU)."""
    g = golden("rollout_cartpole_mlp")
    W, init = g["W"], g["init"]
    tid = [int(i) for i in g["trace_ids"]]
    perm = np.array(tid + [i for i in range(W.shape[0]) if i not in tid])
    P = W.shape[0]
    eng = _engine(population=P, group=P, n_head=1, eval_ep_num=int(g["E"]))
    fit, steps, trace, actions = eng.rollout(0, 0.0, None, w_override=_cuda(W[perm]), init_states=_cuda(init), n_trace=len(tid))
    fit = fit.cpu().numpy()
    # returns: exact for >= 99.9 % of offspring (near-tie action flips vs torch are allowed for the rest)
    assert (fit == g["fitness"][perm]).mean() >= 0.999
    np.testing.assert_allclose(fit, g["fitness"][perm], rtol=1e-4) if (fit == g["fitness"][perm]).all() else None
    trace = trace.cpu().numpy(); actions = actions.cpu().numpy()[:, :, 0]
    for j in range(len(tid)):
        ref = g["traces"][j]
        n = int(np.isfinite(ref[:, 0]).sum())
        assert np.array_equal(actions[j, :n], g["trace_actions"][j, :n])
        assert np.abs(trace[j, :n] - ref[:n]).max() <= 1e-9      # north_star: 1e-9 over the first 200 steps
        # against the bit-twin the trace is exact
        _, ttr, tac = twin.rollout_cartpole(W[perm][j], E=int(g["E"]), init=init, trace_steps=200)
        m = int(np.isfinite(ttr[:, 0]).sum())
        assert np.array_equal(trace[j, :m], ttr[:m]) and np.array_equal(actions[j, :m], tac[:m])


def test_rollout_edge_cases(twin):
    mu = np.zeros((1, D), np.float32)
    # max_step = 1: every episode is truncated after one step
    eng = _engine(population=64, group=64, max_step=1, eval_ep_num=4)
    fit, steps = eng.rollout(0, 1.0, _cuda(mu))
    assert np.all(steps.cpu().numpy() == 4) and np.all(fit.cpu().numpy() == 1.0)
    # the smallest population, the largest E
    eng = _engine(population=2, group=2, eval_ep_num=32, seed=9)
    fit, steps = eng.rollout(5, 2.0, _cuda(mu))
    tf, ts = twin.population_cartpole(mu, sigma=2.0, seed=9, gen=5, group=2, n_head=1, n=2, E=32)
    assert np.array_equal(steps.cpu().numpy(), ts)
    # empty slice: nothing written
    eng = _engine(population=64, group=64, id_begin=10, id_end=10)
    fit, steps = eng.rollout(0, 1.0, _cuda(mu))
    assert np.all(steps.cpu().numpy() == 0)
    # ragged slice writes only its own range
    eng = _engine(population=64, group=64, id_begin=5, id_end=38, seed=2)
    fit, steps = eng.rollout(1, 2.0, _cuda(mu))
    tf, ts = twin.population_cartpole(mu, sigma=2.0, seed=2, gen=1, group=64, n_head=1, id0=5, n=33)
    s = steps.cpu().numpy()
    assert np.array_equal(s[5:38], ts) and np.all(s[:5] == 0) and np.all(s[38:] == 0)


def test_rollout_full_size_properties():
    """BASELINE config 3 size (P = 65536): size-independent properties."""
    P, E = 65536, 5
    mu = _cuda(np.zeros((1, D), np.float32))
    eng = _engine(population=P, group=P, n_head=1, eval_ep_num=E, seed=0)
    fit, steps = eng.rollout(0, 2.0, mu)
    fit2, steps2 = eng.rollout(0, 2.0, mu)
    assert torch.equal(steps, steps2) and torch.equal(fit, fit2)                 # deterministic, schedule independent
    s = steps.cpu().numpy(); f = fit.cpu().numpy()
    assert s.min() >= E * 8 and s.max() <= E * 500 and np.array_equal(f, s / E)
    # sharding: two half-population handles reproduce the unsharded result (no data-path collective)
    a = _engine(population=P, group=P, n_head=1, eval_ep_num=E, seed=0, id_begin=0, id_end=P // 2 + 3)
    b = _engine(population=P, group=P, n_head=1, eval_ep_num=E, seed=0, id_begin=P // 2 + 3, id_end=P)
    fa, sa = a.rollout(0, 2.0, mu)
    fb, sb = b.rollout(0, 2.0, mu, fitness=fa, steps=sa)
    assert torch.equal(sb, steps)
    # offspring 0 is the unperturbed all-zero parent -> action 0 forever
    assert 8 * E <= s[0] <= 12 * E
    # a different generation gives different noise
    _, steps3 = eng.rollout(1, 2.0, mu)
    assert not torch.equal(steps3, steps)


# ------------------------------------------------------------------------------------- K2
@pytest.mark.parametrize("fused", [0, 1, 2])
@pytest.mark.parametrize("n", [2, 97, 4097, 65536, 1 << 20])
def test_rank_desc_bit_exact(n, fused, monkeypatch):
    """All K2 builds: separate init / histogram / scatter / shape kernels (0, test build), the fused build (1: 1 + passes launches)
    and the product's persistent kernel (2: the whole sort in one cooperative launch, grid barriers between the passes)."""
    monkeypatch.setenv("SES_B200_TEST_BUILD", "1")
    monkeypatch.setenv("SES_K2_FUSED", str(min(fused, 1)))
    monkeypatch.setenv("SES_K2_PERSISTENT", "1" if fused == 2 else "0")
    eng = _engine(population=max(n, 2), group=max(n, 2))
    rng = np.random.default_rng(n)
    for kind in ("float", "ties", "cartpole"):
        if kind == "float":
            r = rng.normal(0, 100, n)
        elif kind == "ties":
            r = np.round(rng.normal(0, 3, n))                      # negative values and many ties
        else:
            r = rng.integers(40, 2501, n) / 5.0
        want = np.flip(np.argsort(r, kind="stable")).astype(np.int32)
        l0 = eng.launches
        got = eng.rank_desc(_cuda(r), full_key=True).cpu().numpy()
        assert np.array_equal(got, want), kind
        assert eng.launches - l0 == {0: 18, 1: 9, 2: 1}[fused] - (1 if fused == 0 else 0)      # float64 keys: 8 passes (no shaping kernel here)
        if kind == "cartpole":                                     # integer-key fast path
            got, shaped = eng.rank_desc(_cuda(r), shaped=True)
            assert np.array_equal(got.cpu().numpy(), want)
            cr = ((n - 1 - np.arange(n)) / (n - 1) - 0.5) / np.sqrt((n + 1) / (12.0 * (n - 1)))
            np.testing.assert_allclose(shaped.cpu().numpy()[want], cr, rtol=1e-15, atol=0)


def test_rank_and_shaping_match_reference(twin, golden):
    g = golden("strategy_openai_es")
    P = int(g["P"])
    eng = _engine(population=P, group=P)
    for gen in range(3):
        order, shaped = eng.rank_desc(_cuda(g["rewards_%d" % gen]), shaped=True, full_key=True)
        assert np.array_equal(order.cpu().numpy(), g["order_%d" % gen])            # bit-exact indices
        assert np.array_equal(shaped.cpu().numpy(), twin.centered_rank(g["order_%d" % gen].astype(np.int32)))
        np.testing.assert_allclose(shaped.cpu().numpy(), g["shaped_%d" % gen], rtol=1e-12, atol=1e-15)
    for name in ("strategy_simple_evolution", "strategy_simple_genetic"):
        g = golden(name)
        P, k = int(g["P"]), int(g["cfg_elite_num"])
        eng = _engine(population=P, group=P)
        for gen in range(3):
            order = eng.rank_desc(_cuda(g["rewards_%d" % gen]), full_key=True).cpu().numpy()
            assert np.array_equal(order[:k], g["elite_ids_%d" % gen])              # top-k parents bit-exact


# ------------------------------------------------------------------------------------- K3
def test_update_openai_regenerated_noise_bit_exact(twin):
    P = 4096 + 37
    eng = _engine(population=P, group=P, n_head=1, seed=21)
    rng = np.random.default_rng(2)
    order = rng.permutation(P).astype(np.int32)
    shaped = twin.centered_rank(order)
    mu = rng.normal(0, 1, D).astype(np.float32); m = rng.normal(0, .01, D).astype(np.float32)
    v = np.abs(rng.normal(0, .01, D)).astype(np.float32)
    lr, sigma, t, gen = 0.1, 0.2, 4, 9
    mu_d, m_d, v_d = _cuda(mu), _cuda(m), _cuda(v)
    grad_d = torch.empty(D, dtype=torch.float32, device="cuda")
    eng.update_openai(gen, sigma, lr, t, _cuda(shaped), mu_d, m_d, v_d, grad_out=grad_d)
    g = twin.grad_openai(shaped, D, 21, gen, P, 1, -(lr / (P * sigma)))
    assert np.array_equal(grad_d.cpu().numpy(), g)
    th, mm, vv = twin.adam(mu, m, v, g, eng.adam_a(lr, t))
    assert np.array_equal(mu_d.cpu().numpy(), th) and np.array_equal(m_d.cpu().numpy(), mm) and np.array_equal(v_d.cpu().numpy(), vv)


def test_update_openai_sgd_bit_exact(twin):
    """engine.optimizer: sgd (opt-in; the reference ships Adam only): the fixed-tree gradient followed by
    v = mom*v + (1-mom)*g, theta += -lr*v in float32, bit-exact against the twin (itself pinned on a numpy restatement of the
    OpenAI SGD the reference's optimizers.py names as its source, tests/test_oracle_strategy.py)."""
    P = 4096 + 37
    eng = _engine(population=P, group=P, n_head=1, seed=4)
    rng = np.random.default_rng(6)
    mu = rng.normal(0, 1, D).astype(np.float32); v = rng.normal(0, .01, D).astype(np.float32)
    mu_d, v_d = _cuda(mu), _cuda(v)
    grad_d = torch.empty(D, dtype=torch.float32, device="cuda")
    sigma, lr = 0.3, 0.05
    for gen, mom in [(0, 0.9), (1, 0.9), (2, 0.0)]:
        shaped = twin.centered_rank(rng.permutation(P).astype(np.int32))
        eng.update_openai_sgd(gen, sigma, lr, _cuda(shaped), mu_d, v_d, momentum=mom, grad_out=grad_d)
        g = twin.grad_openai(shaped, D, 4, gen, P, 1, -(lr / (P * sigma)))
        assert np.array_equal(grad_d.cpu().numpy(), g)
        mu, v = twin.sgd(mu, v, g, lr, mom)
        assert np.array_equal(mu_d.cpu().numpy(), mu) and np.array_equal(v_d.cpu().numpy(), v)


def test_update_openai_matches_reference_with_its_noise(golden):
    """Verification mode: the kernel consumes the reference's own epsilon arrays (which hold
    mu+eps from generation 1 on, quirk Q1) and must land on the reference's mu / Adam state."""
    g = golden("strategy_openai_es")
    P, lr = int(g["P"]), float(g["cfg_learning_rate"])
    eng = _engine(population=P, group=P, n_head=1)
    m_d = torch.zeros(D, dtype=torch.float32, device="cuda"); v_d = torch.zeros_like(m_d)
    for gen in range(3):
        sigma = float(g["sigma_before_%d" % gen])
        mu_d = _cuda(g["mu_before_%d" % gen])
        grad_d = torch.empty(D, dtype=torch.float32, device="cuda")
        eng.update_openai(gen, sigma, lr, gen + 1, _cuda(g["shaped_%d" % gen]), mu_d, m_d, v_d,
                          eps_override=_cuda(g["eps_%d" % gen]), grad_out=grad_d)
        np.testing.assert_allclose(grad_d.cpu().numpy(), g["grad_%d" % gen], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(mu_d.cpu().numpy(), g["mu_after_%d" % gen], rtol=1e-4, atol=1e-6)   # north_star rtol
        np.testing.assert_allclose(m_d.cpu().numpy(), g["adam_m_%d" % gen], rtol=1e-4, atol=1e-9)
        np.testing.assert_allclose(v_d.cpu().numpy(), g["adam_v_%d" % gen], rtol=2e-4, atol=1e-12)


def test_elite_mean_and_genetic_carry_over(twin, golden):
    g = golden("strategy_simple_evolution")
    P, k = int(g["P"]), int(g["cfg_elite_num"])
    eng = _engine(population=P, group=P, n_head=2)
    for gen in range(3):                                           # verification mode: reference populations
        order = eng.rank_desc(_cuda(g["rewards_%d" % gen]), full_key=True)
        mu = eng.elite_mean(gen, 0.0, None, order, k, w_override=_cuda(g["pop_%d" % gen]))
        assert np.array_equal(mu.cpu().numpy(), g["mu_after_%d" % gen])              # bit-exact vs the reference
    # the reference's aliased elites (slots 0 and 1 are one module summed onto itself; tests/test_oracle_strategy.py): same bits
    ga = golden("strategy_simple_evolution_alias")
    for gen in range(int(ga["generations"])):
        order = eng.rank_desc(_cuda(ga["rewards_%d" % gen]), full_key=True)
        assert np.array_equal(order.cpu().numpy()[:k], ga["elite_ids_%d" % gen])
        mu = eng.elite_mean(gen, 0.0, None, order, k, w_override=_cuda(ga["pop_%d" % gen]))
        assert np.array_equal(mu.cpu().numpy(), ga["mu_after_%d" % gen])
    # Philox mode vs the twin
    rng = np.random.default_rng(4)
    parent = rng.normal(0, 1, (1, D)).astype(np.float32)
    eng = _engine(population=97, group=97, n_head=2, seed=8)
    order = rng.permutation(97).astype(np.int32)
    mu = eng.elite_mean(3, 1.5, _cuda(parent), _cuda(order), 10).cpu().numpy()
    assert np.array_equal(mu, twin.elite_mean(twin.materialize(parent, 1.5, 8, 3, 97, 2, order[:10])))
    g = golden("strategy_simple_genetic")
    P, k = int(g["P"]), int(g["cfg_elite_num"])
    eng = _engine(population=P, group=P // k, n_head=1, n_parents=k)
    for gen in range(3):
        order = eng.rank_desc(_cuda(g["rewards_%d" % gen]), full_key=True)
        el = eng.materialize(gen, 0.0, None, order[:k].contiguous(), w_override=_cuda(g["pop_%d" % gen]))
        assert np.array_equal(el.cpu().numpy(), g["elites_after_%d" % gen])


def test_generation_openai_host_matches_twin_composition(twin):
    P, E, lr, sigma, seed = 2048, 5, 0.1, 0.5, 17
    eng = _engine(population=P, group=P, n_head=1, eval_ep_num=E, seed=seed)
    mu = np.zeros(D, np.float32); m = np.zeros(D, np.float32); v = np.zeros(D, np.float32)
    fit = np.zeros(P, np.float64)
    tmu, tm, tv = mu.copy(), m.copy(), v.copy()
    for gen in range(3):
        total = eng.generation_openai_host(gen, sigma, lr, gen + 1, mu, m, v, fit)
        tf, ts = twin.population_cartpole(tmu[None], sigma=sigma, seed=seed, gen=gen, group=P, n_head=1, n=P, E=E, nthreads=8)
        assert total == ts.sum() and np.array_equal(fit, tf)
        shaped = twin.centered_rank(twin.rank_desc(tf))
        g = twin.grad_openai(shaped, D, seed, gen, P, 1, -(lr / (P * sigma)))
        tmu, tm, tv = twin.adam(tmu, tm, tv, g, eng.adam_a(lr, gen + 1))
        assert np.array_equal(mu, tmu) and np.array_equal(m, tm) and np.array_equal(v, tv)
        sigma *= 0.999


# ------------------------------------------------------------------------------------- block-cyclic sharding
@pytest.mark.parametrize("env,obs,act,gru,P,world", [("CartPole-v1", 4, 2, False, 3001, 2), ("CartPole-v1", 4, 2, True, 301, 3),
                                                     ("simple_spread", 12, 5, False, 1000, 8), ("Acrobot-v1", 6, 3, False, 777, 2)])
def test_block_cyclic_shards_reproduce_the_whole_population(twin, env, obs, act, gru, P, world):
    """The ranks' block-cyclic slices (engine.owned_ids / Shard::local_to_id) together give exactly the fitness vector
    of one engine rolling out everything; verification-mode weights and traces are addressed in local order."""
    from simple_es_b200.engine import cyclic_block, owned_ids
    kw = dict(env_name=env, obs_dim=obs, act_dim=act, gru=gru, population=P, group=P, n_head=1, eval_ep_num=3, seed=8,
              max_step=None, init_mode="fresh")
    whole = _engine(**kw)
    rng = np.random.default_rng(2)
    mu = rng.normal(0, 0.4, (1, whole.D)).astype(np.float32)
    fit1, steps1 = whole.rollout(2, 0.9, _cuda(mu))
    B = cyclic_block(P, world)
    fit = torch.full((P,), float("nan"), dtype=torch.float64, device="cuda"); steps = torch.full((P,), -1, dtype=torch.int64, device="cuda")
    for r in range(world):
        eng = _engine(shard=(r, world, B), **kw)
        assert eng.n_local == owned_ids(P, r, world, B).size
        eng.rollout(2, 0.9, _cuda(mu), fitness=fit, steps=steps)
        eng.close()
    assert torch.equal(fit, fit1) and torch.equal(steps, steps1)
    # verification mode on rank 1's slice: explicit weights row l belong to offspring owned_ids[l]
    ids = owned_ids(P, 1, world, B)
    W = whole.materialize(2, 0.9, _cuda(mu), _cuda(ids.astype(np.int32)))
    eng = _engine(shard=(1, world, B), **kw)
    f2 = torch.zeros(P, dtype=torch.float64, device="cuda"); s2 = torch.zeros(P, dtype=torch.int64, device="cuda")
    out = eng.rollout(2, 0.0, None, fitness=f2, steps=s2, w_override=W.contiguous(), n_trace=min(4, ids.size))
    assert torch.equal(f2[torch.from_numpy(ids).cuda()], fit1[torch.from_numpy(ids).cuda()])


# ------------------------------------------------------------------------------------- opt-in: antithetic sampling
@pytest.mark.parametrize("strategy,n,k,gru", [("openai_es", 1025, None, False), ("simple_genetic", 1000, 8, False), ("simple_evolution", 300, 10, True)])
def test_antithetic_sampling_bit_exact_and_mirrored(twin, strategy, n, k, gru):
    """engine.antithetic (not in the reference): the perturbed offspring of a group come in (+eps, -eps) pairs.  K1, the
    materialiser and the openai_es gradient re-derive the same mirrored noise as the oracle, bit for bit; with a zero
    parent the two members of a pair are exact negatives of each other."""
    from simple_es_b200.engine import population_layout
    P, group, n_head, n_par = population_layout(strategy, n, k)
    Dn = 6562 if gru else D
    eng = _engine(population=P, group=group, n_head=n_head, n_parents=n_par, gru=gru, seed=17, antithetic=True)
    rng = np.random.default_rng(1)
    parents = rng.normal(0, 0.3, (n_par, Dn)).astype(np.float32)
    twin.set_antithetic(True)
    try:
        ids = np.arange(P, dtype=np.int32)
        got = eng.materialize(5, 0.7, _cuda(parents), _cuda(ids)).cpu().numpy()
        want = twin.materialize(parents, 0.7, 17, 5, group, n_head, ids)
        assert np.array_equal(got, want)
        zero = eng.materialize(5, 0.7, _cuda(np.zeros_like(parents)), _cuda(ids)).cpu().numpy()
        for g0 in range(0, P, group):
            pert = zero[g0 + n_head:g0 + group]
            m = (pert.shape[0] // 2) * 2
            assert np.array_equal(pert[0:m:2], -pert[1:m:2]) and np.any(pert[0] != 0)
        fit, steps = eng.rollout(5, 0.7, _cuda(parents))
        tf, ts = twin.population_cartpole(parents, gru=gru, sigma=0.7, seed=17, gen=5, group=group, n_head=n_head, n=P, E=5, nthreads=8)
        assert np.array_equal(steps.cpu().numpy(), ts) and np.array_equal(fit.cpu().numpy(), tf)
        if strategy == "openai_es":
            order, shaped = eng.rank_desc(fit, shaped=True)
            mu = _cuda(parents[0].copy()); m_ = torch.zeros(Dn, device="cuda"); v_ = torch.zeros(Dn, device="cuda")
            gout = torch.zeros(Dn, device="cuda")
            eng.update_openai(5, 0.7, 0.1, 1, shaped, mu, m_, v_, grad_out=gout)
            g = twin.grad_openai(twin.centered_rank(twin.rank_desc(tf)), Dn, 17, 5, group, n_head, -(0.1 / (P * 0.7)))
            assert np.array_equal(gout.cpu().numpy(), g)
    finally:
        twin.set_antithetic(False)
    # and the switch really changes the population
    plain = _engine(population=P, group=group, n_head=n_head, n_parents=n_par, gru=gru, seed=17)
    assert not np.array_equal(plain.materialize(5, 0.7, _cuda(parents), _cuda(ids)).cpu().numpy(), got)


# ------------------------------------------------------------------------------------- K1: GRU policy (BASELINE config 2)
DG = 6562


@pytest.mark.parametrize("pomdp,E,sigma", [(True, 5, 0.7), (False, 3, 0.3), (True, 7, 0.7)])
def test_rollout_gru_philox_bit_exact(twin, pomdp, E, sigma):
    P = 600
    eng = _engine(population=P, group=P, n_head=2, eval_ep_num=E, gru=True, pomdp=pomdp, seed=13)
    rng = np.random.default_rng(7)
    mu = rng.normal(0, 0.3, (1, DG)).astype(np.float32)
    fit, steps = eng.rollout(4, sigma, _cuda(mu))
    tf, ts = twin.population_cartpole(mu, gru=True, pomdp=pomdp, sigma=sigma, seed=13, gen=4, group=P, n_head=2, n=P, E=E, nthreads=8)
    assert np.array_equal(steps.cpu().numpy(), ts)
    assert np.array_equal(fit.cpu().numpy(), tf)
    assert ts[0] == ts[1]                                   # simple_evolution layout: offspring 0 and 1 are both mu


def test_rollout_gru_verification_mode_matches_reference(twin, golden):
    g = golden("rollout_cartpole_gru_pomdp")
    W, init, E = g["W"], g["init"], int(g["E"])
    tid = [int(i) for i in g["trace_ids"]]
    perm = np.array(tid + [i for i in range(W.shape[0]) if i not in tid])
    P = W.shape[0]
    eng = _engine(population=P, group=P, n_head=1, eval_ep_num=E, gru=True, pomdp=True)
    fit, steps, trace, actions = eng.rollout(0, 0.0, None, w_override=_cuda(W[perm]), init_states=_cuda(init), n_trace=len(tid))
    assert (fit.cpu().numpy() == g["fitness"][perm]).mean() >= 0.999
    trace = trace.cpu().numpy(); actions = actions.cpu().numpy()[:, :, 0]
    for j in range(len(tid)):
        ref = g["traces"][j]
        n = int(np.isfinite(ref[:, 0]).sum())
        assert np.array_equal(actions[j, :n], g["trace_actions"][j, :n])
        assert np.abs(trace[j, :n] - ref[:n]).max() <= 1e-9


# ------------------------------------------------------------------------------------- K1: simple_spread (BASELINE config 4)
@pytest.mark.parametrize("N,E,init_mode", [(2, 5, "shared"), (3, 5, "shared"), (2, 4, "fresh"), (3, 2, "fresh")])
def test_rollout_spread_philox_bit_exact(twin, N, E, init_mode):
    P = 1500
    Dn = 6 * N * 32 + 32 + 5 * 32 + 5
    eng = _engine(env_name="simple_spread", obs_dim=6 * N, act_dim=5, n_agents=N, max_step="None", population=P, group=P,
                  n_head=1, eval_ep_num=E, seed=19, init_mode=init_mode)
    assert eng.D == Dn
    rng = np.random.default_rng(3)
    mu = rng.normal(0, 0.5, (1, Dn)).astype(np.float32)
    fit, steps = eng.rollout(6, 0.8, _cuda(mu))
    tf, ts = twin.population_mpe(mu, N=N, sigma=0.8, seed=19, gen=6, group=P, n_head=1, n=P, E=E,
                                 init_mode=0 if init_mode == "shared" else 1)
    assert np.array_equal(steps.cpu().numpy(), ts) and np.all(ts == 25 * E)
    assert np.array_equal(fit.cpu().numpy(), tf)            # float64 returns under the contract: bit-exact


@pytest.mark.parametrize("name", ["rollout_spread_n2", "rollout_spread_n3"])
def test_rollout_spread_verification_mode_matches_reference(twin, golden, name):
    """Golden = reference RolloutWorker + GymEnvModel copies per agent over the float64 numpy restatement of
    simple_spread.  Returns must agree to rtol 1e-4 (north_star); they agree to ~1e-15 (the reference keeps one
    running sum over all episodes, the engine sums per episode first)."""
    g = golden(name)
    W, init, N, E = g["W"], g["init"], int(g["N"]), int(g["E"])
    P = W.shape[0]
    eng = _engine(env_name="simple_spread", obs_dim=6 * N, act_dim=5, n_agents=N, max_step="None", population=P, group=P,
                  n_head=1, eval_ep_num=E)
    nt = g["traces"].shape[0]
    fit, steps, trace, actions = eng.rollout(0, 0.0, None, w_override=_cuda(W), init_states=_cuda(init), n_trace=nt)
    np.testing.assert_allclose(fit.cpu().numpy(), g["fitness"], rtol=1e-12)
    trace = trace.cpu().numpy(); actions = actions.cpu().numpy()
    assert np.array_equal(actions[:, :25], g["trace_actions"])
    assert np.abs(trace[:, :25] - g["traces"]).max() <= 1e-9


# ------------------------------------------------------------------------------------- BASELINE full sizes: size-independent properties
def test_full_size_gru_config_properties(twin):
    """BASELINE config 1: CartPole POMDP + GRU, simple_evolution, offspring_num 4096 (P = 4097)."""
    P, E = 4097, 5
    rng = np.random.default_rng(5)
    mu = _cuda(rng.normal(0, 0.3, (1, DG)).astype(np.float32))
    eng = _engine(population=P, group=P, n_head=2, eval_ep_num=E, gru=True, pomdp=True, seed=2)
    fit, steps = eng.rollout(3, 0.5, mu)
    fit2, steps2 = eng.rollout(3, 0.5, mu)
    assert torch.equal(steps, steps2)                                             # deterministic
    s = steps.cpu().numpy()
    assert s[0] == s[1] and s.min() >= 5 * 8 and s.max() <= 5 * 500 and np.array_equal(fit.cpu().numpy(), s / E)
    a = _engine(population=P, group=P, n_head=2, eval_ep_num=E, gru=True, pomdp=True, seed=2, id_begin=0, id_end=1500)
    b = _engine(population=P, group=P, n_head=2, eval_ep_num=E, gru=True, pomdp=True, seed=2, id_begin=1500, id_end=P)
    fa, sa = a.rollout(3, 0.5, mu)
    b.rollout(3, 0.5, mu, fitness=fa, steps=sa)
    assert torch.equal(sa, steps)                                                 # shards reproduce the whole
    ids = rng.choice(P, 64, replace=False)                                        # spot-check against the oracle
    for i in ids:
        w = twin.materialize(mu.cpu().numpy(), 0.5, 2, 3, P, 2, np.array([i], np.int32))[0]
        t, _, _ = twin.rollout_cartpole(w, gru=True, pomdp=True, E=E, seed=2, gen=3, idx=int(i))
        assert t == s[i]


def test_full_size_spread_config_properties(twin):
    """BASELINE config 3: simple_spread, shared MLP, openai_es, population 16384 (N = 2 as the reference, and N = 3)."""
    for N in (2, 3):
        P, E = 16384, 5
        Dn = 6 * N * 32 + 32 + 165
        eng = _engine(env_name="simple_spread", obs_dim=6 * N, act_dim=5, n_agents=N, max_step="None", population=P, group=P,
                      n_head=1, eval_ep_num=E, seed=4, init_mode="fresh")
        mu = _cuda(np.zeros((1, Dn), np.float32))
        fit, steps = eng.rollout(0, 0.2, mu)
        fit2, _ = eng.rollout(0, 0.2, mu)
        assert torch.equal(fit, fit2) and torch.all(steps == 25 * E)
        f = fit.cpu().numpy()
        assert np.all(f < 0) and np.isfinite(f).all()                             # rewards are negative distances / collisions
        order, shaped = eng.rank_desc(fit, shaped=True)
        o = order.cpu().numpy()
        assert np.array_equal(np.sort(o), np.arange(P)) and np.all(np.diff(f[o]) <= 0)      # a permutation, descending
        sh = shaped.cpu().numpy()
        assert abs(sh.mean()) < 1e-12 and abs(sh.std() - 1) < 1e-12 and sh[o[0]] == sh.max()
        for i in np.random.default_rng(N).choice(P, 48, replace=False):
            w = twin.materialize(np.zeros((1, Dn), np.float32), 0.2, 4, 0, P, 1, np.array([i], np.int32))[0]
            tf, _, _, _ = twin.rollout_mpe(w, N=N, E=E, seed=4, init_mode=1, gen=0, idx=int(i))
            assert tf == f[i]


def test_full_size_genetic_config_properties():
    """BASELINE config 4: CartPole simple_genetic, k*(n//k) = 2^20 offspring, 16 elites."""
    P, k = 1 << 20, 16
    eng = _engine(population=P, group=P // k, n_head=1, n_parents=k, seed=1)
    rng = np.random.default_rng(0)
    parents = _cuda(rng.normal(0, 1, (k, D)).astype(np.float32))
    fit, steps = eng.rollout(2, 1.0, parents)
    s = steps.cpu().numpy()
    assert s.min() >= 40 and s.max() <= 2500
    order = eng.rank_desc(fit).cpu().numpy()
    want = np.flip(np.argsort(fit.cpu().numpy(), kind="stable"))
    assert np.array_equal(order, want)                                            # bit-exact top-k / full order at 2^20 with heavy ties
    elites = eng.materialize(2, 1.0, parents, _cuda(order[:k].astype(np.int32))).cpu().numpy()
    heads = np.arange(k) * (P // k)                                               # first of each group is the unperturbed elite
    fit_heads = fit.cpu().numpy()[heads]
    el_heads = eng.materialize(2, 1.0, parents, _cuda(heads.astype(np.int32))).cpu().numpy()
    assert np.array_equal(el_heads, parents.cpu().numpy()) and np.isfinite(elites).all() and fit_heads.shape == (k,)
