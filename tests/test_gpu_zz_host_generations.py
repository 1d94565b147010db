"""GPU: the host-buffer whole-generation entry points of the two elite strategies (ses_generation_evolution_host,
ses_generation_genetic_host -- one C call per ESLoop.run iteration, loop.py:61-84, numpy buffers in and out) against the
same generations composed from the CPU twin.  Bit-exact.  (This file sorts last on purpose: these entry points were added at
the very end of round 1, after the GPU budget was spent -- they are verified on the host SIMT emulator,
tests/test_simt_emu.py -- so a surprise here cannot mask the rest of the -m gpu suite under `-x`.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

D = 226


def _engine(strategy, n, k, **kw):
    from simple_es_b200.engine import RolloutEngine, population_layout
    P, group, n_head, n_par = population_layout(strategy, n, k)
    return RolloutEngine("CartPole-v1", 4, 2, False, False, 500, 5, P, group, n_head, n_par, **kw), (P, group, n_head, n_par)


def test_generation_evolution_host_matches_twin_composition(twin):
    seed = 31
    eng, (P, group, n_head, n_par) = _engine("simple_evolution", 2000, 10, seed=seed)
    mu = np.zeros(D, np.float32); fit = np.zeros(P); tmu = mu.copy(); sigma = 2.0
    for gen in range(3):
        total = eng.generation_evolution_host(gen, sigma, 10, mu, fit)
        tf, ts = twin.population_cartpole(tmu[None], sigma=sigma, seed=seed, gen=gen, group=group, n_head=n_head, n=P, E=5, nthreads=8)
        order = twin.rank_desc(tf)
        tmu = twin.elite_mean(twin.materialize(tmu[None], sigma, seed, gen, group, n_head, order[:10]))
        assert total == ts.sum() and np.array_equal(fit, tf) and np.array_equal(mu, tmu)
        sigma *= 0.9


def test_generation_genetic_host_matches_twin_composition(twin):
    seed = 32
    eng, (P, group, n_head, n_par) = _engine("simple_genetic", 4096, 8, seed=seed)
    el = np.random.default_rng(1).normal(0, 0.5, (n_par, D)).astype(np.float32); fit = np.zeros(P); tel = el.copy(); sigma = 1.0
    for gen in range(3):
        total = eng.generation_genetic_host(gen, sigma, el, fit)
        tf, ts = twin.population_cartpole(tel, sigma=sigma, seed=seed, gen=gen, group=group, n_head=n_head, n=P, E=5, nthreads=8)
        tel = twin.materialize(tel, sigma, seed, gen, group, n_head, twin.rank_desc(tf)[:n_par])
        assert total == ts.sum() and np.array_equal(fit, tf) and np.array_equal(el, tel)
        sigma *= 0.95


def test_generation_elite_hosts_reject_bad_handles():
    eng, (P, group, n_head, n_par) = _engine("simple_genetic", 400, 4, seed=1)
    with pytest.raises(RuntimeError, match="one parent"):
        eng.generation_evolution_host(0, 1.0, 3, np.zeros(D, np.float32), np.zeros(P))
    from simple_es_b200.engine import RolloutEngine
    sliced = RolloutEngine("CartPole-v1", 4, 2, False, False, 500, 5, P, group, n_head, n_par, id_begin=0, id_end=50)
    with pytest.raises(RuntimeError, match="single-slice"):
        sliced.generation_genetic_host(0, 1.0, np.zeros((n_par, D), np.float32), np.zeros(P))


@pytest.mark.parametrize("variant", [6, 7])
def test_k1_variants_6_7_branch_free_division_bit_exact(twin, monkeypatch, variant):
    """K1 variants 6 / 7 (opt-in, SES_K1_VARIANT): the CartPole step with a branch-free double division (7), and with the
    action-dependent tail evaluated for both actions as well (6).  Same bits as the twin; the division equals IEEE division
    on 2^30 random in-range operand pairs."""
    from simple_es_b200.engine import RolloutEngine
    monkeypatch.setenv("SES_B200_TEST_BUILD", "1")
    monkeypatch.setenv("SES_K1_VARIANT", str(variant))
    eng = RolloutEngine("CartPole-v1", 4, 2, False, False, 500, 5, 3000, 3000, 1, 1, seed=11)
    assert eng.test_ddiv_fast(1 << 30) == 0
    mu = np.zeros((1, D), np.float32)
    fit, steps = eng.rollout(2, 2.0, torch.from_numpy(mu).cuda())
    tf, ts = twin.population_cartpole(mu, sigma=2.0, seed=11, gen=2, group=3000, n_head=1, n=3000, E=5, nthreads=8)
    assert np.array_equal(steps.cpu().numpy(), ts) and np.array_equal(fit.cpu().numpy(), tf)
    w1 = mu[0, :128].reshape(32, 4); w2 = mu[0, 160:224].reshape(2, 32)
    w1[0] = [0.0, 0.5, 10.0, 3.0]; w2[1, 0] = 5.0; w2[0, 0] = -5.0       # a policy that balances: 500-step episodes
    fit, steps = eng.rollout(0, 0.05, torch.from_numpy(mu).cuda())
    tf, ts = twin.population_cartpole(mu, sigma=0.05, seed=11, gen=0, group=3000, n_head=1, n=3000, E=5, nthreads=8)
    assert np.array_equal(steps.cpu().numpy(), ts) and (ts == 2500).mean() > 0.5


@pytest.mark.parametrize("variant", [0, 1, 2, 4])
def test_gru_test_build_variants_bit_exact(twin, monkeypatch, variant):
    """SES_GRU_VARIANT (test build): 0 the plain GRU rollout (physics after the argmax), 1 speculative physics with every table in
    shared memory, 2 a warp pair per offspring over named barriers (3 + 2 episodes), 4 two tables in registers -- the product's
    kernel (3: speculative physics, n-gate table in registers) is what every other GRU test runs.  Same bits as the twin
    (E = 5 and a chunked E = 7, POMDP on / off)."""
    from simple_es_b200.engine import RolloutEngine
    monkeypatch.setenv("SES_B200_TEST_BUILD", "1")
    monkeypatch.setenv("SES_GRU_VARIANT", str(variant))
    rng = np.random.default_rng(7)
    mu = rng.normal(0, 0.3, (1, 6562)).astype(np.float32)
    for pomdp, E in [(True, 5), (False, 7)]:
        P = 600
        eng = RolloutEngine("CartPole-v1", 4, 2, True, pomdp, 500, E, P, P, 2, 1, seed=13)
        fit, steps = eng.rollout(4, 0.7, torch.from_numpy(mu).cuda())
        tf, ts = twin.population_cartpole(mu, gru=True, pomdp=pomdp, sigma=0.7, seed=13, gen=4, group=P, n_head=2, n=P, E=E, nthreads=8)
        assert np.array_equal(steps.cpu().numpy(), ts) and np.array_equal(fit.cpu().numpy(), tf)
