"""The engine's CUDA sources, compiled for the HOST on the SIMT emulator of tests/simt_emu (every CUDA thread a fiber;
see tests/simt_emu/include/simt_emu.h), against the CPU oracle and the reference-generated golden vectors.

Purpose: the dev container has no GPU, so without this the kernels' logic (warp schedulers, ballots / shuffles /
match_any, shared-memory layouts, Philox counters, operation order of the numerical contract) could only be checked on
the B200 box.  These tests run the SAME kernel source through the SAME C ABI entry points (ses_abi.cu) at small sizes and
demand what the GPU parity tests (tests/test_gpu_*.py, -m gpu) demand at full size: bit equality with the oracle.
The emulated library is test infrastructure only -- the product never loads it and has no CPU fallback
(test_host_logic.py::test_no_cpu_fallback); performance, inter-warp races and the memory model remain GPU-only questions
(compute-sanitizer: profiles/r01_sanitizer.txt)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

D = 226
DG = 6562


@pytest.fixture(scope="module")
def emu():
    from simt_emu import emu_engine
    emu_engine.load()
    return emu_engine.EmuEngine


def _trained_mu():
    """A parent that balances the pole (bang-bang on theta + theta_dot): 500-step episodes with small sigma."""
    mu = np.zeros((1, D), np.float32)
    w1 = mu[0, :128].reshape(32, 4); w2 = mu[0, 160:224].reshape(2, 32)
    w1[0] = [0.0, 0.5, 10.0, 3.0]; w2[1, 0] = 5.0; w2[0, 0] = -5.0
    return mu


# ------------------------------------------------------------------------------------- contract
def test_emu_math_contract_bit_exact(emu, twin):
    eng = emu()
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-10, 10, 200_000), rng.normal(0, 1, 200_000), [0.0, -0.0, 50, -50]]).astype(np.float32)
    assert np.array_equal(eng.test_math("tanh", x), twin.tanhf(x))
    assert np.array_equal(eng.test_math("sigmoid", x), twin.sigmf(x))
    assert np.array_equal(eng.test_math("tanh_fast", x), twin.tanhf(x))
    # packed / scalar fast-path tanh == the contract's tanh on a slice of the float32 line (the GPU test walks all of it;
    # here the reciprocal seed is the correctly rounded 1/q instead of MUFU.RCP, so this checks the algebra, not the seed)
    assert eng.test_tanh_x2_exhaustive(False, 0.25, 0.2500305) == 0
    assert eng.test_tanh_x2_exhaustive(True, 3.0, 3.0002441) == 0
    # K1 variant 6's branch-free double division (Newton / Markstein from a 20-bit reciprocal seed) == IEEE division
    assert eng.test_ddiv_fast(3_000_000) == 0
    u = ((rng.integers(0, 2 ** 24, 200_000) + 0.5) * 2.0 ** -24).astype(np.float32)
    assert np.array_equal(eng.test_math("ln", u), twin.lnf(u))
    v = (rng.integers(0, 2 ** 24, 200_000) * 2.0 ** -24).astype(np.float32)
    s, c = twin.sincos2pif(v)
    assert np.array_equal(eng.test_math("sin2pi", v), s) and np.array_equal(eng.test_math("cos2pi", v), c)
    th = rng.uniform(-0.5, 0.5, 200_000)
    s, c = twin.sincos(th)
    assert np.array_equal(eng.test_math("sin64", th), s) and np.array_equal(eng.test_math("cos64", th), c)
    xf = np.concatenate([rng.uniform(-100, 100, 200_000), np.arange(-80, 81) * (np.pi / 4), [0.0, -0.0]])
    s, c = twin.sincos_full(xf)
    assert np.array_equal(eng.test_math("sin64_full", xf), s) and np.array_equal(eng.test_math("cos64_full", xf), c)


def test_emu_philox_normals_bit_exact(emu, twin):
    eng = emu(seed=1234)
    for gen, idx in [(0, 0), (0, 1), (3, 77), (1000, 65535), (2 ** 31, 2 ** 20 - 1)]:
        assert np.array_equal(eng.test_normals(gen, idx), twin.normals(1234, gen, idx, D))


@pytest.mark.parametrize("strategy,n,k", [("simple_evolution", 96, 10), ("openai_es", 128, None), ("simple_genetic", 100, 8)])
def test_emu_materialize_layouts_bit_exact(emu, twin, strategy, n, k):
    from simple_es_b200.engine import population_layout
    P, group, n_head, n_par = population_layout(strategy, n, k)
    eng = emu(population=P, group=group, n_head=n_head, n_parents=n_par, seed=5)
    parents = np.random.default_rng(1).normal(0, 1, (n_par, D)).astype(np.float32)
    ids = np.arange(P, dtype=np.int32)
    got = eng.materialize(7, 0.37, parents, ids)
    assert np.array_equal(got, twin.materialize(parents, 0.37, 5, 7, group, n_head, ids))
    assert all(np.array_equal(got[i], parents[i // group]) for i in range(P) if i % group < n_head)


# ------------------------------------------------------------------------------------- K1: CartPole MLP (slot kernel)
@pytest.mark.parametrize("init_mode,pomdp,E", [("shared", False, 5), ("fresh", False, 3), ("shared", True, 5), ("shared", False, 1),
                                               ("fresh", False, 7)])
def test_emu_rollout_philox_bit_exact(emu, twin, init_mode, pomdp, E):
    P = 333
    eng = emu(population=P, group=P, n_head=1, eval_ep_num=E, pomdp=pomdp, init_mode=init_mode, seed=11)
    mu = np.zeros((1, D), np.float32)
    fit, steps = eng.rollout(2, 2.0, mu)
    tf, ts = twin.population_cartpole(mu, pomdp=pomdp, sigma=2.0, seed=11, gen=2, group=P, n_head=1, n=P, E=E,
                                      init_mode=0 if init_mode == "shared" else 1, nthreads=4)
    assert np.array_equal(steps, ts) and np.array_equal(fit, tf)
    assert ts.max() > 3 * ts.min()                          # ragged episode lengths: the warp scheduler re-arms lanes


@pytest.mark.parametrize("knobs", [{}, {"SES_K1_TAIL": "0"}, {"SES_K1_TAIL": "1"}, {"SES_K1_TAIL": "100"}, {"SES_K1_SPLIT": "0"},
                                   {"SES_K1_SPARSE_RANK": "0", "SES_K1_SPARSE_QUOTA": "3"}, {"SES_K1_SPARSE_RANK": "1", "SES_K1_SPARSE_QUOTA": "7"}])
def test_emu_rollout_launch_geometry_never_changes_a_bit(emu, twin, knobs, monkeypatch):
    """The round-2 scheduler (exact-request queue with offspring shared between warps, straggler phase, sparse warps): fitness
    and step counts do not depend on the launch geometry -- 500-step and ragged populations, twice in a row (the cross-warp
    accumulators return to zero)."""
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    for P, E, sigma, mu in [(90, 5, 0.02, _trained_mu()), (150, 5, 2.0, np.zeros((1, D), np.float32)), (61, 3, 0.02, _trained_mu())]:
        eng = emu(population=P, group=P, n_head=1, eval_ep_num=E, seed=31, max_step=120)
        for gen in (0, 1):
            fit, steps = eng.rollout(gen, sigma, mu)
            tf, ts = twin.population_cartpole(mu, sigma=sigma, seed=31, gen=gen, group=P, n_head=1, n=P, E=E, max_step=120, nthreads=4)
            assert np.array_equal(steps, ts) and np.array_equal(fit, tf)


@pytest.mark.parametrize("sms,ctas", [(2, 3), (3, 2)])
def test_emu_rollout_multi_round_converged_launches(emu, twin, sms, ctas, monkeypatch):
    """Multi-round launches of full-length episodes (queue A, then the exact queue B for the last two rounds' worth), three
    generations each so that the adaptive tail sees a converged previous launch; 2 / 3 resident CTAs per emulated SM.  The
    launcher's choice is read back through the geometry hook; same bits as the twin."""
    monkeypatch.setenv("SES_SIMT_EMU_SMS", str(sms))
    monkeypatch.setenv("SES_SIMT_EMU_CTAS_PER_SM", str(ctas))
    for P in (230, 270, 420):
        eng = emu(population=P, group=P, n_head=1, eval_ep_num=5, seed=37, max_step=60)
        for gen in (0, 1, 2):
            fit, steps = eng.rollout(gen, 0.02, _trained_mu())
            g = eng.test_k1_geometry()
            assert g["ctas_per_sm"] == ctas and g["resident_warps"] == sms * ctas * 4 and g["lanes"] == 32 and g["sparse_rank"] < 0
            assert g["tail_start"] == max(0, (P * 5 - 2 * 32 * g["resident_warps"]) // 5 * 5)       # two rounds of exact requests
            tf, ts = twin.population_cartpole(_trained_mu(), sigma=0.02, seed=37, gen=gen, group=P, n_head=1, n=P, E=5, max_step=60, nthreads=4)
            assert np.array_equal(steps, ts) and np.array_equal(fit, tf)
            assert (ts == 300).mean() > 0.9                 # converged: (nearly) every episode runs to the truncation
        eng.close()


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5, 6, 7])
def test_emu_rollout_k1_variants_bit_exact(emu, twin, variant, monkeypatch):
    """Every K1 code path (scalar / packed FFMA2, permuted slot table, weights of the lane's slot in registers) is the
    same function: ragged and 500-step episodes, Philox and verification (w_override) inputs."""
    monkeypatch.setenv("SES_K1_VARIANT", str(variant))
    for P, sigma, seed, pomdp, mu in [(160, 2.0, 21, False, np.zeros((1, D), np.float32)), (40, 0.05, 22, False, _trained_mu()),
                                      (96, 1.0, 23, True, np.zeros((1, D), np.float32))]:
        eng = emu(population=P, group=P, n_head=1, eval_ep_num=5, seed=seed, pomdp=pomdp)
        fit, steps = eng.rollout(1, sigma, mu)
        tf, ts = twin.population_cartpole(mu, pomdp=pomdp, sigma=sigma, seed=seed, gen=1, group=P, n_head=1, n=P, E=5, nthreads=4)
        assert np.array_equal(steps, ts) and np.array_equal(fit, tf)
        if sigma < 1.0:
            assert (ts == 2500).mean() > 0.5               # most episodes hit the 500-step truncation
        eng.close()
    rng = np.random.default_rng(5)
    W = rng.normal(0, 1.5, (64, D)).astype(np.float32)
    init = rng.uniform(-0.05, 0.05, (5, 4))
    eng = emu(population=64, group=64, n_head=1, eval_ep_num=5)
    fit, steps = eng.rollout(0, 0.0, None, w_override=W, init_states=init)
    tf, ts = twin.population_cartpole(np.zeros((1, D), np.float32), n=64, group=64, E=5, W_override=W, init=init, nthreads=4)
    assert np.array_equal(steps, ts) and np.array_equal(fit, tf)


@pytest.mark.parametrize("case", range(8))
def test_emu_rollout_scheduler_random_shapes(emu, twin, case, monkeypatch):
    """The warp scheduler's result does not depend on how work is laid out: random population sizes, episode counts, lanes
    taking work (SES_ROLLOUT_LANES), grid limits (SES_ROLLOUT_CTAS_PER_SM, emulated SM count) and slices -- same bits."""
    rng = np.random.default_rng(100 + case)
    E = int(rng.choice([1, 2, 3, 4, 5, 6, 8, 11, 32]))
    P = int(rng.integers(2, 260))
    lo = int(rng.integers(0, P)); hi = int(rng.integers(lo, P + 1))
    if rng.random() < 0.5:
        monkeypatch.setenv("SES_ROLLOUT_LANES", str(int(rng.integers(1, 33))))
    monkeypatch.setenv("SES_ROLLOUT_CTAS_PER_SM", str(int(rng.integers(1, 3))))
    monkeypatch.setenv("SES_SIMT_EMU_SMS", str(int(rng.integers(1, 4))))
    init_mode = "fresh" if rng.random() < 0.5 else "shared"
    n_head = int(rng.integers(0, 3))
    eng = emu(population=P, group=P, n_head=n_head, eval_ep_num=E, seed=case, init_mode=init_mode, id_begin=lo, id_end=hi,
              max_step=int(rng.choice([500, 37])))
    mu = rng.normal(0, 0.2, (1, D)).astype(np.float32)
    fit, steps = eng.rollout(case, 1.0, mu)
    tf, ts = twin.population_cartpole(mu, sigma=1.0, seed=case, gen=case, group=P, n_head=n_head, id0=lo, n=hi - lo, E=E,
                                      max_step=eng.max_step, init_mode=0 if init_mode == "shared" else 1, nthreads=2)
    assert np.array_equal(steps[lo:hi], ts) and np.array_equal(fit[lo:hi], tf)
    assert np.all(steps[:lo] == -1) and np.all(steps[hi:] == -1)


def test_emu_spread_eight_slot_build(emu, twin, monkeypatch):
    """SES_SPREAD_SLOTS8=1 selects the 8-slot simple_spread instantiations (the default for E >= 5 is 6 slots): same bits."""
    monkeypatch.setenv("SES_SPREAD_SLOTS8", "1")
    for N in (2, 3):
        P, Dn = 70, 6 * N * 32 + 32 + 165
        eng = emu(env_name="simple_spread", obs_dim=6 * N, act_dim=5, n_agents=N, max_step="None", population=P, group=P,
                  n_head=1, eval_ep_num=5, seed=5)
        mu = np.random.default_rng(N).normal(0, 0.5, (1, Dn)).astype(np.float32)
        fit, steps = eng.rollout(1, 0.6, mu)
        tf, ts = twin.population_mpe(mu, N=N, sigma=0.6, seed=5, gen=1, group=P, n_head=1, n=P, E=5)
        assert np.array_equal(steps, ts) and np.array_equal(fit, tf)


@pytest.mark.parametrize("E", [3, 1])
def test_emu_rollout_16_and_32_slots_per_warp(emu, twin, E):
    P = 200
    eng = emu(population=P, group=P, n_head=1, eval_ep_num=E, seed=31)
    mu = np.zeros((1, D), np.float32)
    fit, steps = eng.rollout(4, 1.5, mu)
    tf, ts = twin.population_cartpole(mu, sigma=1.5, seed=31, gen=4, group=P, n_head=1, n=P, E=E, nthreads=4)
    assert np.array_equal(steps, ts) and np.array_equal(fit, tf)


def test_emu_rollout_verification_mode_matches_reference(emu, twin, golden):
    """The kernel source consumes the reference's own weight arrays and initial states (north_star verification mode).
    Golden = reference RolloutWorker + GymEnvModel (torch CPU) over the float64 CartPole restatement."""
    g = golden("rollout_cartpole_mlp")
    W, init = g["W"], g["init"]
    tid = [int(i) for i in g["trace_ids"]]
    perm = np.array(tid + [i for i in range(W.shape[0]) if i not in tid])
    P = W.shape[0]
    eng = emu(population=P, group=P, n_head=1, eval_ep_num=int(g["E"]))
    fit, steps, trace, actions = eng.rollout(0, 0.0, None, w_override=W[perm], init_states=init, n_trace=len(tid))
    assert (fit == g["fitness"][perm]).mean() >= 0.999
    actions = actions[:, :, 0]
    for j in range(len(tid)):
        ref = g["traces"][j]
        n = int(np.isfinite(ref[:, 0]).sum())
        assert np.array_equal(actions[j, :n], g["trace_actions"][j, :n])
        assert np.abs(trace[j, :n] - ref[:n]).max() <= 1e-9      # north_star: 1e-9 over the first 200 steps
        _, ttr, tac = twin.rollout_cartpole(W[perm][j], E=int(g["E"]), init=init, trace_steps=200)
        m = int(np.isfinite(ttr[:, 0]).sum())
        assert np.array_equal(trace[j, :m], ttr[:m]) and np.array_equal(actions[j, :m], tac[:m])


def test_emu_rollout_edge_cases(emu, twin):
    mu = np.zeros((1, D), np.float32)
    eng = emu(population=64, group=64, max_step=1, eval_ep_num=4)            # every episode truncated after one step
    fit, steps = eng.rollout(0, 1.0, mu)
    assert np.all(steps == 4) and np.all(fit == 1.0)
    eng = emu(population=2, group=2, eval_ep_num=32, seed=9)                 # smallest population, largest E
    fit, steps = eng.rollout(5, 2.0, mu)
    tf, ts = twin.population_cartpole(mu, sigma=2.0, seed=9, gen=5, group=2, n_head=1, n=2, E=32)
    assert np.array_equal(steps, ts)
    eng = emu(population=64, group=64, id_begin=10, id_end=10)               # empty slice: nothing written
    fit, steps = eng.rollout(0, 1.0, mu)
    assert np.all(steps == -1) and np.isnan(fit).all()
    eng = emu(population=64, group=64, id_begin=5, id_end=38, seed=2)        # ragged slice writes only its own range
    fit, steps = eng.rollout(1, 2.0, mu)
    tf, ts = twin.population_cartpole(mu, sigma=2.0, seed=2, gen=1, group=64, n_head=1, id0=5, n=33)
    assert np.array_equal(steps[5:38], ts) and np.all(steps[:5] == -1) and np.all(steps[38:] == -1)
    eng = emu(population=50, group=50, seed=3)                               # CartPole-v0 = the 200-step limit
    v0 = emu(env_name="CartPole-v0", population=50, group=50, seed=3, max_step=None)
    _, s500 = eng.rollout(0, 0.05, _trained_mu())
    _, s200 = v0.rollout(0, 0.05, _trained_mu())
    assert s200.max() == 1000 and np.array_equal(s200, np.minimum(s200, 1000)) and s500.max() == 2500


def test_emu_engine_rejects_bad_configs(emu):
    with pytest.raises(RuntimeError, match="num_state=4"):
        emu(obs_dim=5)
    with pytest.raises(RuntimeError, match="eval_ep_num"):
        emu(eval_ep_num=33)
    with pytest.raises(RuntimeError, match="population"):
        emu(population=1, group=1)
    with pytest.raises(RuntimeError, match="parents"):
        emu(population=64, group=16, n_parents=3)
    with pytest.raises(RuntimeError, match="N=2 or N=3"):
        emu(env_name="simple_spread", obs_dim=24, act_dim=5, n_agents=4)


# ------------------------------------------------------------------------------------- K2
@pytest.mark.parametrize("fused", [0, 1])
@pytest.mark.parametrize("n", [2, 97, 4097, 65536 + 3, (1 << 18) + 5])
def test_emu_rank_desc_bit_exact(emu, n, fused, monkeypatch):
    """Both K2 builds: separate init / histogram / scatter / shape kernels (2 + 2*passes launches) and the fused one
    (SES_K2_FUSED=1: 1 + passes launches) give the same permutation -- np.flip(np.argsort(kind="stable"))."""
    monkeypatch.setenv("SES_K2_FUSED", str(fused))
    eng = emu(population=n, group=n)
    rng = np.random.default_rng(n)
    kinds = ("float", "ties", "cartpole") if n <= 4097 else ("cartpole",)
    for kind in kinds:
        if kind == "float":
            r = rng.normal(0, 100, n)
        elif kind == "ties":
            r = np.round(rng.normal(0, 3, n))                      # negative values, -0.0 / +0.0 and many ties
            r[::7] = -0.0
        else:
            r = rng.integers(40, 2501, n) / 5.0
        want = np.flip(np.argsort(r, kind="stable")).astype(np.int32)
        if n <= 4097:
            assert np.array_equal(eng.rank_desc(r, full_key=True), want), kind
        if kind == "cartpole":                                     # integer-key fast path (2 radix passes)
            l0 = eng.launches
            order, shaped = eng.rank_desc(r, shaped=True)
            assert np.array_equal(order, want) and eng.launches - l0 == (3 if fused else 6)
            cr = ((n - 1 - np.arange(n)) / (n - 1) - 0.5) / np.sqrt((n + 1) / (12.0 * (n - 1)))
            np.testing.assert_allclose(shaped[want], cr, rtol=1e-15, atol=0)
            neg = emu(env_name="MountainCar-v0", obs_dim=2, act_dim=3, population=n, group=n, max_step=None)
            assert np.array_equal(neg.rank_desc(-r), np.flip(np.argsort(-r, kind="stable")).astype(np.int32))


@pytest.mark.parametrize("fused", [0, 1])
def test_emu_rank_and_shaping_match_reference(emu, twin, golden, fused, monkeypatch):
    monkeypatch.setenv("SES_K2_FUSED", str(fused))
    g = golden("strategy_openai_es")
    P = int(g["P"])
    eng = emu(population=P, group=P)
    for gen in range(3):
        order, shaped = eng.rank_desc(g["rewards_%d" % gen], shaped=True, full_key=True)
        assert np.array_equal(order, g["order_%d" % gen])                          # bit-exact indices
        assert np.array_equal(shaped, twin.centered_rank(g["order_%d" % gen].astype(np.int32)))
        np.testing.assert_allclose(shaped, g["shaped_%d" % gen], rtol=1e-12, atol=1e-15)
    for name in ("strategy_simple_evolution", "strategy_simple_genetic"):
        g = golden(name)
        P, k = int(g["P"]), int(g["cfg_elite_num"])
        eng = emu(population=P, group=P)
        for gen in range(3):
            assert np.array_equal(eng.rank_desc(g["rewards_%d" % gen], full_key=True)[:k], g["elite_ids_%d" % gen])


# ------------------------------------------------------------------------------------- K3
def test_emu_update_openai_regenerated_noise_bit_exact(emu, twin):
    P = 4096 + 37                                                   # 130 level-0 blocks -> 3 level-1 groups (one ragged)
    eng = emu(population=P, group=P, n_head=1, seed=21)
    rng = np.random.default_rng(2)
    shaped = twin.centered_rank(rng.permutation(P).astype(np.int32))
    mu = rng.normal(0, 1, D).astype(np.float32); m = rng.normal(0, .01, D).astype(np.float32)
    v = np.abs(rng.normal(0, .01, D)).astype(np.float32)
    lr, sigma, t, gen = 0.1, 0.2, 4, 9
    mu_d, m_d, v_d = mu.copy(), m.copy(), v.copy()
    grad = eng.update_openai(gen, sigma, lr, t, shaped, mu_d, m_d, v_d)
    g = twin.grad_openai(shaped, D, 21, gen, P, 1, -(lr / (P * sigma)))
    assert np.array_equal(grad, g)
    th, mm, vv = twin.adam(mu, m, v, g, eng.adam_a(lr, t))
    assert np.array_equal(mu_d, th) and np.array_equal(m_d, mm) and np.array_equal(v_d, vv)


def test_emu_update_openai_sgd_bit_exact(emu, twin):
    """engine.optimizer: sgd (opt-in) -- same fixed-tree gradient, then v = mom*v + (1-mom)*g, theta += -lr*v in float32."""
    P = 1000
    eng = emu(population=P, group=P, n_head=1, seed=4)
    rng = np.random.default_rng(6)
    mu = rng.normal(0, 1, D).astype(np.float32); v = rng.normal(0, .01, D).astype(np.float32)
    tmu, tv = mu.copy(), v.copy()
    sigma, lr = 0.3, 0.05
    for gen, mom in [(0, 0.9), (1, 0.9), (2, 0.0)]:
        shaped = twin.centered_rank(rng.permutation(P).astype(np.int32))
        grad = eng.update_openai_sgd(gen, sigma, lr, shaped, mu, v, momentum=mom)
        g = twin.grad_openai(shaped, D, 4, gen, P, 1, -(lr / (P * sigma)))
        assert np.array_equal(grad, g)
        tmu, tv = twin.sgd(tmu, tv, g, lr, mom)
        assert np.array_equal(mu, tmu) and np.array_equal(v, tv)
    with pytest.raises(RuntimeError, match="momentum"):
        eng.update_openai_sgd(0, sigma, lr, shaped, mu, v, momentum=1.0)


def test_emu_update_openai_matches_reference_with_its_noise(emu, golden):
    g = golden("strategy_openai_es")
    P, lr = int(g["P"]), float(g["cfg_learning_rate"])
    eng = emu(population=P, group=P, n_head=1)
    m_d = np.zeros(D, np.float32); v_d = np.zeros(D, np.float32)
    for gen in range(3):
        sigma = float(g["sigma_before_%d" % gen])
        mu_d = g["mu_before_%d" % gen].astype(np.float32).copy()
        grad = eng.update_openai(gen, sigma, lr, gen + 1, g["shaped_%d" % gen], mu_d, m_d, v_d, eps_override=g["eps_%d" % gen])
        np.testing.assert_allclose(grad, g["grad_%d" % gen], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(mu_d, g["mu_after_%d" % gen], rtol=1e-4, atol=1e-6)      # north_star rtol
        np.testing.assert_allclose(m_d, g["adam_m_%d" % gen], rtol=1e-4, atol=1e-9)
        np.testing.assert_allclose(v_d, g["adam_v_%d" % gen], rtol=2e-4, atol=1e-12)


def test_emu_elite_mean_and_genetic_carry_over(emu, twin, golden):
    g = golden("strategy_simple_evolution")
    P, k = int(g["P"]), int(g["cfg_elite_num"])
    eng = emu(population=P, group=P, n_head=2)
    for gen in range(3):                                           # verification mode: the reference's populations
        order = eng.rank_desc(g["rewards_%d" % gen], full_key=True)
        assert np.array_equal(eng.elite_mean(gen, 0.0, None, order, k, w_override=g["pop_%d" % gen]), g["mu_after_%d" % gen])
    ga = golden("strategy_simple_evolution_alias")                 # the reference's aliased elite slots: same bits
    for gen in range(int(ga["generations"])):
        order = eng.rank_desc(ga["rewards_%d" % gen], full_key=True)
        assert np.array_equal(eng.elite_mean(gen, 0.0, None, order, k, w_override=ga["pop_%d" % gen]), ga["mu_after_%d" % gen])
    rng = np.random.default_rng(4)
    parent = rng.normal(0, 1, (1, D)).astype(np.float32)
    eng = emu(population=97, group=97, n_head=2, seed=8)
    order = rng.permutation(97).astype(np.int32)
    assert np.array_equal(eng.elite_mean(3, 1.5, parent, order, 10),
                          twin.elite_mean(twin.materialize(parent, 1.5, 8, 3, 97, 2, order[:10])))
    g = golden("strategy_simple_genetic")
    P, k = int(g["P"]), int(g["cfg_elite_num"])
    eng = emu(population=P, group=P // k, n_head=1, n_parents=k)
    for gen in range(3):
        order = eng.rank_desc(g["rewards_%d" % gen], full_key=True)
        assert np.array_equal(eng.materialize(gen, 0.0, None, order[:k], w_override=g["pop_%d" % gen]), g["elites_after_%d" % gen])


@pytest.mark.parametrize("fused", [0, 1])
def test_emu_generation_openai_host_matches_twin_composition(emu, twin, fused, monkeypatch):
    """ses_generation_openai_host (bench.py's e2e entry point): K1 -> K2 (integer keys) -> K3 composed inside the library."""
    monkeypatch.setenv("SES_K2_FUSED", str(fused))
    P, E, lr, sigma, seed = 300, 5, 0.1, 0.5, 17
    eng = emu(population=P, group=P, n_head=1, eval_ep_num=E, seed=seed)
    mu = np.zeros(D, np.float32); m = np.zeros(D, np.float32); v = np.zeros(D, np.float32)
    fit = np.zeros(P, np.float64)
    tmu, tm, tv = mu.copy(), m.copy(), v.copy()
    for gen in range(3):
        total = eng.generation_openai_host(gen, sigma, lr, gen + 1, mu, m, v, fit)
        tf, ts = twin.population_cartpole(tmu[None], sigma=sigma, seed=seed, gen=gen, group=P, n_head=1, n=P, E=E, nthreads=4)
        assert total == ts.sum() and np.array_equal(fit, tf)
        shaped = twin.centered_rank(twin.rank_desc(tf))
        g = twin.grad_openai(shaped, D, seed, gen, P, 1, -(lr / (P * sigma)))
        tmu, tm, tv = twin.adam(tmu, tm, tv, g, eng.adam_a(lr, gen + 1))
        assert np.array_equal(mu, tmu) and np.array_equal(m, tm) and np.array_equal(v, tv)
        sigma *= 0.999
    # K1 + K2 (sort init + 2 x (hist, scatter) + shape, or fused: first hist + 2 scatters) + 2 x K3
    assert eng.launches == 3 * (6 if fused else 9)


def test_emu_generation_elite_hosts_match_twin_composition(emu, twin):
    """ses_generation_evolution_host / ses_generation_genetic_host: one C call per generation with host buffers, equal to
    rollout -> rank -> elite mean / elite carry-over composed from the twin (incl. each strategy's sigma-decay ordering)."""
    from simple_es_b200.engine import population_layout
    E, seed = 5, 31
    # simple_evolution: offspring_num 60 -> P = 61, layout [mu, mu, 59 perturbed]
    P, group, n_head, n_par = population_layout("simple_evolution", 60, 7)
    eng = emu(population=P, group=group, n_head=n_head, n_parents=n_par, eval_ep_num=E, seed=seed)
    mu = np.zeros(D, np.float32); fit = np.zeros(P); tmu = mu.copy(); sigma = 2.0
    for gen in range(3):
        total = eng.generation_evolution_host(gen, sigma, 7, mu, fit)
        tf, ts = twin.population_cartpole(tmu[None], sigma=sigma, seed=seed, gen=gen, group=group, n_head=n_head, n=P, E=E, nthreads=4)
        order = twin.rank_desc(tf)
        tmu = twin.elite_mean(twin.materialize(tmu[None], sigma, seed, gen, group, n_head, order[:7]))
        assert total == ts.sum() and np.array_equal(fit, tf) and np.array_equal(mu, tmu)
        sigma *= 0.9
    # simple_genetic: 4 elites x 25 offspring each
    P, group, n_head, n_par = population_layout("simple_genetic", 100, 4)
    eng = emu(population=P, group=group, n_head=n_head, n_parents=n_par, eval_ep_num=E, seed=seed)
    el = np.random.default_rng(1).normal(0, 0.5, (n_par, D)).astype(np.float32); fit = np.zeros(P); tel = el.copy(); sigma = 1.0
    for gen in range(3):
        total = eng.generation_genetic_host(gen, sigma, el, fit)
        tf, ts = twin.population_cartpole(tel, sigma=sigma, seed=seed, gen=gen, group=group, n_head=n_head, n=P, E=E, nthreads=4)
        tel = twin.materialize(tel, sigma, seed, gen, group, n_head, twin.rank_desc(tf)[:n_par])
        assert total == ts.sum() and np.array_equal(fit, tf) and np.array_equal(el, tel)
        sigma *= 0.95
    with pytest.raises(RuntimeError, match="one parent"):
        eng.generation_evolution_host(0, 1.0, 3, el[0].copy(), fit)
    sliced = emu(population=P, group=group, n_head=n_head, n_parents=n_par, eval_ep_num=E, seed=seed, id_begin=0, id_end=50)
    with pytest.raises(RuntimeError, match="single-slice"):
        sliced.generation_genetic_host(0, 1.0, el, fit)


# ------------------------------------------------------------------------------------- sharding, antithetic
@pytest.mark.parametrize("env,obs,act,gru,P,world", [("CartPole-v1", 4, 2, False, 301, 2), ("CartPole-v1", 4, 2, True, 41, 3),
                                                     ("simple_spread", 12, 5, False, 100, 8), ("Acrobot-v1", 6, 3, False, 77, 2)])
def test_emu_block_cyclic_shards_reproduce_the_whole_population(emu, env, obs, act, gru, P, world):
    from simple_es_b200.engine import cyclic_block, owned_ids
    kw = dict(env_name=env, obs_dim=obs, act_dim=act, gru=gru, population=P, group=P, n_head=1, eval_ep_num=3, seed=8,
              max_step=60 if env == "Acrobot-v1" else None, init_mode="fresh")
    whole = emu(**kw)
    mu = np.random.default_rng(2).normal(0, 0.4, (1, whole.D)).astype(np.float32)
    fit1, steps1 = whole.rollout(2, 0.9, mu)
    B = cyclic_block(P, world)
    fit = np.full(P, np.nan); steps = np.full(P, -1, dtype=np.int64)
    for r in range(world):
        eng = emu(shard=(r, world, B), **kw)
        assert eng.n_local == owned_ids(P, r, world, B).size
        eng.rollout(2, 0.9, mu, fitness=fit, steps=steps)
        eng.close()
    assert np.array_equal(fit, fit1) and np.array_equal(steps, steps1)
    ids = owned_ids(P, 1, world, B)                                # verification mode: rows in local (work-queue) order
    W = whole.materialize(2, 0.9, mu, ids.astype(np.int32))
    eng = emu(shard=(1, world, B), **kw)
    f2 = np.zeros(P); s2 = np.zeros(P, dtype=np.int64)
    eng.rollout(2, 0.0, None, fitness=f2, steps=s2, w_override=W, n_trace=min(4, ids.size))
    assert np.array_equal(f2[ids], fit1[ids])


@pytest.mark.parametrize("strategy,n,k,gru", [("openai_es", 257, None, False), ("simple_genetic", 120, 8, False), ("simple_evolution", 24, 10, True)])
def test_emu_antithetic_sampling_bit_exact_and_mirrored(emu, twin, strategy, n, k, gru):
    from simple_es_b200.engine import population_layout
    P, group, n_head, n_par = population_layout(strategy, n, k)
    Dn = DG if gru else D
    eng = emu(population=P, group=group, n_head=n_head, n_parents=n_par, gru=gru, seed=17, antithetic=True)
    parents = np.random.default_rng(1).normal(0, 0.3, (n_par, Dn)).astype(np.float32)
    twin.set_antithetic(True)
    try:
        ids = np.arange(P, dtype=np.int32)
        got = eng.materialize(5, 0.7, parents, ids)
        assert np.array_equal(got, twin.materialize(parents, 0.7, 17, 5, group, n_head, ids))
        zero = eng.materialize(5, 0.7, np.zeros_like(parents), ids)
        for g0 in range(0, P, group):
            pert = zero[g0 + n_head:g0 + group]
            m = (pert.shape[0] // 2) * 2
            assert np.array_equal(pert[0:m:2], -pert[1:m:2]) and np.any(pert[0] != 0)
        fit, steps = eng.rollout(5, 0.7, parents)
        tf, ts = twin.population_cartpole(parents, gru=gru, sigma=0.7, seed=17, gen=5, group=group, n_head=n_head, n=P, E=5, nthreads=4)
        assert np.array_equal(steps, ts) and np.array_equal(fit, tf)
        if strategy == "openai_es":
            order, shaped = eng.rank_desc(fit, shaped=True)
            mu = parents[0].copy(); m_ = np.zeros(Dn, np.float32); v_ = np.zeros(Dn, np.float32)
            gout = eng.update_openai(5, 0.7, 0.1, 1, shaped, mu, m_, v_)
            g = twin.grad_openai(twin.centered_rank(twin.rank_desc(tf)), Dn, 17, 5, group, n_head, -(0.1 / (P * 0.7)))
            assert np.array_equal(gout, g)
    finally:
        twin.set_antithetic(False)


# ------------------------------------------------------------------------------------- K1: GRU policy (warp per offspring)
@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("pomdp,E,sigma", [(True, 5, 0.7), (False, 3, 0.3), (True, 7, 0.7), (False, 1, 0.5), (False, 2, 0.5), (True, 4, 0.6),
                                           (False, 11, 0.4)])
def test_emu_rollout_gru_philox_bit_exact(emu, twin, pomdp, E, sigma, variant, monkeypatch):
    """SES_GRU_VARIANT (test build) 0: plain single-warp kernel; 1: the physics evaluated for both actions on otherwise idle lanes
    at the start of the step; 2 (the product's kernel): a warp PAIR per offspring, episodes split 3 + 2 (E = 1 falls back to 1)."""
    monkeypatch.setenv("SES_GRU_VARIANT", str(variant))
    P = 40
    eng = emu(population=P, group=P, n_head=2, eval_ep_num=E, gru=True, pomdp=pomdp, seed=13)
    mu = np.random.default_rng(7).normal(0, 0.3, (1, DG)).astype(np.float32)
    fit, steps = eng.rollout(4, sigma, mu)
    tf, ts = twin.population_cartpole(mu, gru=True, pomdp=pomdp, sigma=sigma, seed=13, gen=4, group=P, n_head=2, n=P, E=E, nthreads=4)
    assert np.array_equal(steps, ts) and np.array_equal(fit, tf)
    assert ts[0] == ts[1]                                   # simple_evolution layout: offspring 0 and 1 are both mu


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_emu_rollout_gru_verification_mode_matches_reference(emu, golden, variant, monkeypatch):
    monkeypatch.setenv("SES_GRU_VARIANT", str(variant))
    g = golden("rollout_cartpole_gru_pomdp")
    W, init, E = g["W"], g["init"], int(g["E"])
    tid = [int(i) for i in g["trace_ids"]]
    perm = np.array(tid + [i for i in range(W.shape[0]) if i not in tid])
    P = W.shape[0]
    eng = emu(population=P, group=P, n_head=1, eval_ep_num=E, gru=True, pomdp=True)
    fit, steps, trace, actions = eng.rollout(0, 0.0, None, w_override=W[perm], init_states=init, n_trace=len(tid))
    assert (fit == g["fitness"][perm]).mean() >= 0.999
    actions = actions[:, :, 0]
    for j in range(len(tid)):
        ref = g["traces"][j]
        n = int(np.isfinite(ref[:, 0]).sum())
        assert np.array_equal(actions[j, :n], g["trace_actions"][j, :n])
        assert np.abs(trace[j, :n] - ref[:n]).max() <= 1e-9


# ------------------------------------------------------------------------------------- K1: simple_spread
@pytest.mark.parametrize("N,E,init_mode", [(2, 5, "shared"), (3, 5, "shared"), (2, 4, "fresh"), (3, 2, "fresh")])
def test_emu_rollout_spread_philox_bit_exact(emu, twin, N, E, init_mode):
    P = 150
    Dn = 6 * N * 32 + 32 + 5 * 32 + 5
    eng = emu(env_name="simple_spread", obs_dim=6 * N, act_dim=5, n_agents=N, max_step="None", population=P, group=P,
              n_head=1, eval_ep_num=E, seed=19, init_mode=init_mode)
    assert eng.D == Dn
    mu = np.random.default_rng(3).normal(0, 0.5, (1, Dn)).astype(np.float32)
    fit, steps = eng.rollout(6, 0.8, mu)
    tf, ts = twin.population_mpe(mu, N=N, sigma=0.8, seed=19, gen=6, group=P, n_head=1, n=P, E=E,
                                 init_mode=0 if init_mode == "shared" else 1)
    assert np.array_equal(steps, ts) and np.all(ts == 25 * E)
    assert np.array_equal(fit, tf)                          # float64 returns under the contract: bit-exact


@pytest.mark.parametrize("name", ["rollout_spread_n2", "rollout_spread_n3"])
def test_emu_rollout_spread_verification_mode_matches_reference(emu, golden, name):
    g = golden(name)
    W, init, N, E = g["W"], g["init"], int(g["N"]), int(g["E"])
    P = W.shape[0]
    eng = emu(env_name="simple_spread", obs_dim=6 * N, act_dim=5, n_agents=N, max_step="None", population=P, group=P,
              n_head=1, eval_ep_num=E)
    nt = g["traces"].shape[0]
    fit, steps, trace, actions = eng.rollout(0, 0.0, None, w_override=W, init_states=init, n_trace=nt)
    np.testing.assert_allclose(fit, g["fitness"], rtol=1e-12)
    assert np.array_equal(actions[:, :25], g["trace_actions"])
    assert np.abs(trace[:, :25] - g["traces"]).max() <= 1e-9


# ------------------------------------------------------------------------------------- K1: MountainCar-v0, Acrobot-v1
CLASSIC = {"MountainCar-v0": (2, 3, 195, "rollout_mountaincar"), "Acrobot-v1": (6, 3, 323, "rollout_acrobot")}


@pytest.mark.parametrize("env", list(CLASSIC))
@pytest.mark.parametrize("E,init_mode,sigma", [(5, "shared", 2.0), (3, "fresh", 1.0), (1, "fresh", 3.0)])
def test_emu_rollout_classic_philox_bit_exact(emu, twin, env, E, init_mode, sigma):
    P = 64 if env == "Acrobot-v1" else 120
    obs, act, Dn, _ = CLASSIC[env]
    eng = emu(env_name=env, obs_dim=obs, act_dim=act, max_step=None, population=P, group=P, eval_ep_num=E, seed=23, init_mode=init_mode)
    assert eng.D == Dn
    mu = np.random.default_rng(4).normal(0, 0.5, (1, Dn)).astype(np.float32)
    fit, steps = eng.rollout(3, sigma, mu)
    tf, ts = twin.population_classic(env, mu, sigma=sigma, seed=23, gen=3, group=P, n_head=1, n=P, E=E,
                                     init_mode=0 if init_mode == "shared" else 1)
    assert np.array_equal(steps, ts) and np.array_equal(fit, tf)


@pytest.mark.parametrize("env", list(CLASSIC))
def test_emu_rollout_classic_traces_equal_the_twin(emu, twin, golden, env):
    """Verification mode on the reference-driven golden inputs: the emulated kernel's traces are the twin's, bit for bit
    (how far the twin is from the reference's libm path is the subject of tests/test_oracle_classic.py)."""
    obs, act, Dn, name = CLASSIC[env]
    g = golden(name)
    W, init, E = g["W"][:12], g["init"], int(g["E"])
    eng = emu(env_name=env, obs_dim=obs, act_dim=act, max_step=None, population=12, group=12, eval_ep_num=E)
    fit, steps, trace, actions = eng.rollout(0, 0.0, None, w_override=W, init_states=init, n_trace=12)
    for j in range(12):
        tf, tsteps, ttr, tac = twin.rollout_classic(env, W[j], E=E, init=init, trace_steps=200)
        m = int(np.isfinite(ttr[:, 0]).sum())
        assert m > 0 and np.array_equal(trace[j, :m], ttr[:m]) and np.array_equal(actions[j, :m, 0], tac[:m])
        assert fit[j] == tf and steps[j] == tsteps


# ------------------------------------------------------------------------------------- K1: Pendulum-v0, continuous-action head
@pytest.mark.parametrize("E,init_mode,sigma", [(5, "shared", 1.0), (3, "fresh", 0.5), (1, "fresh", 2.0)])
def test_emu_rollout_pendulum_philox_bit_exact(emu, twin, E, init_mode, sigma):
    """The tanh head (networks/neural_network.py:32-33) + Pendulum-v0 in the slot kernel == the twin, bit for bit (float64 returns)."""
    P = 96
    eng = emu(env_name="Pendulum-v0", obs_dim=3, act_dim=1, max_step=200, population=P, group=P, eval_ep_num=E, seed=29, init_mode=init_mode)
    assert eng.D == 161
    mu = np.random.default_rng(6).normal(0, 0.5, (1, 161)).astype(np.float32)
    fit, steps = eng.rollout(4, sigma, mu)
    tf, ts = twin.population_classic("Pendulum-v0", mu, sigma=sigma, seed=29, gen=4, group=P, n_head=1, n=P, E=E,
                                     init_mode=0 if init_mode == "shared" else 1)
    assert np.array_equal(steps, ts) and np.all(ts == 200 * E) and np.array_equal(fit, tf)


def test_emu_rollout_pendulum_traces_equal_the_twin(emu, twin, golden):
    g = golden("rollout_pendulum")
    W, init, E = g["W"][:10], g["init"], int(g["E"])
    eng = emu(env_name="Pendulum-v0", obs_dim=3, act_dim=1, max_step=200, population=10, group=10, eval_ep_num=E)
    fit, steps, trace, actions = eng.rollout(0, 0.0, None, w_override=W, init_states=init, n_trace=10)
    for j in range(10):
        tf, tsteps, ttr, tac = twin.rollout_classic("Pendulum-v0", W[j], E=E, init=init, trace_steps=200)
        assert np.array_equal(trace[j], ttr) and np.array_equal(actions[j, :, 0].view(np.float32), tac)      # float32 action bit patterns
        assert fit[j] == tf and steps[j] == tsteps
    np.testing.assert_allclose(fit, g["fitness"][:10], rtol=1e-4)            # vs the reference path (north_star tolerance)


def test_emu_create_rejects_more_than_2_30_episodes(emu):
    """The episode queue is 32-bit: population x eval_ep_num is bounded at create time (the reference has no such scale)."""
    from simple_es_b200 import _lib as product_lib
    import ctypes as C
    eng = emu()
    P = (1 << 26) + 8
    cfg = product_lib.ses_config(env=0, obs_dim=4, act_dim=2, eval_ep_num=32, population=P, group=P, n_head=1, n_parents=1, id_end=P)
    h = C.c_void_p()
    assert eng.lib.ses_create(C.byref(cfg), C.byref(h)) != 0 and b"2^30 episodes" in eng.lib.ses_last_error()


def test_emu_continuous_head_needs_pendulum(emu):
    from simple_es_b200 import _lib as product_lib
    import ctypes as C
    eng = emu()
    cfg = product_lib.ses_config(env=0, obs_dim=4, act_dim=2, eval_ep_num=5, population=8, group=8, n_head=1, n_parents=1, id_end=8,
                                 continuous_action=1)
    h = C.c_void_p()
    assert eng.lib.ses_create(C.byref(cfg), C.byref(h)) != 0 and b"continuous-action head" in eng.lib.ses_last_error()


# ------------------------------------------------------------------------------------- multi-GPU: peer exchange fused into K1
def _openai_rank(emu, rank, world, P, gens, seed, barrier, handles, out, shard_mode):
    """One emulated rank (an OS thread): the steps of strategies.OpenAIES.step() with the peer exchange."""
    from simple_es_b200.engine import cyclic_block, shard_bounds
    try:
        if world == 1:
            eng = emu(population=P, group=P, n_head=1, seed=seed)
            bufs = [np.zeros(P), np.zeros(P)]
        else:
            if shard_mode == "cyclic":
                eng = emu(population=P, group=P, n_head=1, seed=seed, shard=(rank, world, cyclic_block(P, world)))
            else:
                lo, hi = shard_bounds(P, rank, world)
                eng = emu(population=P, group=P, n_head=1, seed=seed, id_begin=lo, id_end=hi)
            handles[rank] = eng.peer_export()
            barrier.wait()
            bufs = eng.peer_attach(handles, rank, world)
            barrier.wait()
        mu = np.zeros((1, D), np.float32); m = np.zeros(D, np.float32); v = np.zeros(D, np.float32)
        steps = np.zeros(P, dtype=np.int64)
        sigma, lr, hist = 0.5, 0.1, []
        for gen in range(gens):
            fit = bufs[gen & 1]
            eng.rollout(gen, sigma, mu, fitness=fit, steps=steps)        # K1 stores each fitness into every peer's buffer
            if world > 1:
                eng.peer_barrier()
            order, shaped = eng.rank_desc(fit, shaped=True)
            eng.update_openai(gen, sigma, lr, gen + 1, shaped, mu[0], m, v)   # gradient rows sharded over the ranks + barrier
            sigma *= 0.999
            hist.append((fit.copy(), order.copy(), mu.copy(), m.copy(), v.copy()))
        if world > 1:
            eng.peer_check()
            barrier.wait()                                               # nobody frees its buffer while a peer may still write
        out[rank] = hist
    except BaseException as exc:                                         # pragma: no cover
        out[rank] = exc
        barrier.abort()
        raise


@pytest.mark.parametrize("world,shard_mode,P,fold", [(2, "cyclic", 301, 1), (2, "cyclic", 301, 0), (3, "contiguous", 200, 1), (8, "cyclic", 2200, 1)])
def test_emu_peer_exchange_ranks_equal_single_rank(emu, world, shard_mode, P, fold, monkeypatch):
    """SURVEY 8e on the emulator: W ranks (threads), each rolling out its shard with the fitness exchange fused into K1
    (stores into every peer's buffer + flag barrier, double buffered by generation parity) and its share of the gradient's
    level-1 rows, finish every generation with the SAME fitness vector, rank order and (mu, m, v) as one rank doing it all.
    fold = 0 (the product): separate barrier kernels; fold = 1 (SES_PEER_FOLD=1, test build: measured on 8 B200s and not adopted):
    the flag barriers run in a sentinel CTA of K1 and in the last CTA of k_grad_partial (PeerSync).  P = 301 at W = 2 leaves
    rank 0 without gradient rows: it meets the folded barrier of rank 1 with a separate one."""
    import threading
    monkeypatch.setenv("SES_PEER_FOLD", str(fold))
    gens, seed = 3, 29
    ref = {}
    _openai_rank(emu, 0, 1, P, gens, seed, None, None, ref, None)
    barrier = threading.Barrier(world, timeout=120)
    handles, out = [None] * world, {}
    threads = [threading.Thread(target=_openai_rank, args=(emu, r, world, P, gens, seed, barrier, handles, out, shard_mode))
               for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    for r in range(world):
        assert not isinstance(out.get(r), BaseException) and out.get(r) is not None, out.get(r)
        for gen in range(gens):
            for a, b in zip(out[r][gen], ref[0][gen]):
                assert np.array_equal(a, b), (r, gen)


def test_emu_fuzz_all_envs_against_twin():
    """A slice of tools/emu_fuzz.py (random env / policy / layout / slice / E / truncation / antithetic / grid limits):
    emulated kernels == twin, bit for bit."""
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "emu_fuzz.py"), "20000", "40"], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "cases 40 bad 0" in out.stdout, out.stdout[-2000:]
