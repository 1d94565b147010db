"""simple_spread: the bit-twin against numpy's logaddexp and against the reference-driven golden
vectors (reference RolloutWorker + GymEnvModel over oracle/pyref.py::SimpleSpreadShim)."""
import numpy as np
import pytest


def test_logaddexp0_matches_numpy(twin):
    rng = np.random.default_rng(0)
    y = np.concatenate([rng.uniform(-690, 690, 100_000), rng.uniform(-5, 5, 100_000), [0.0, -700.0, 700.0]])
    got, want = twin.logaddexp0(y), np.logaddexp(0, y)
    assert np.all(np.abs(got - want) <= 4 * np.spacing(want) + 1e-300)
    assert np.all(twin.logaddexp0(np.array([-701.0, -1e4])) == 0.0)        # flushed tail (< 1e-304)


@pytest.mark.parametrize("name", ["rollout_spread_n2", "rollout_spread_n3"])
def test_spread_rollout_golden(twin, golden, name):
    g = golden(name)
    N, E, W = int(g["N"]), int(g["E"]), g["W"]
    fit, steps = twin.population_mpe(np.zeros((1, W.shape[1]), np.float32), N=N, n=W.shape[0], E=E, W_override=W, init=g["init"])
    assert np.all(steps == 25 * E)                                           # max_cycles = 25, wrapper max_step "None"
    np.testing.assert_allclose(fit, g["fitness"], rtol=1e-12)                # north_star asks rtol 1e-4
    for i in range(g["traces"].shape[0]):
        f, s, tr, ac = twin.rollout_mpe(W[i], N=N, E=E, init=g["init"], trace_steps=25)
        assert np.array_equal(ac, g["trace_actions"][i])
        assert np.abs(tr - g["traces"][i]).max() <= 1e-9


def test_spread_init_and_zero_policy(twin):
    s = twin.spread_init(3, 0, 0, 0, 0, N=3)
    assert s.shape == (12,) and np.all(np.abs(s) <= 1.0)
    assert np.array_equal(s, twin.spread_init(3, 0, 7, 9, 0, N=3))           # shared table ignores (gen, id)
    assert not np.array_equal(twin.spread_init(3, 1, 0, 1, 0, N=3), twin.spread_init(3, 1, 0, 2, 0, N=3))
    # all-zero weights -> logits 0 -> action 0 (no-op) for every agent: nobody moves
    D = 581
    f, n, tr, ac = twin.rollout_mpe(np.zeros(D, np.float32), N=2, E=1, trace_steps=25)
    assert np.all(ac == 0) and n == 25
    assert np.all(tr[:, 4:] == 0) and np.all(tr[:, :4] == tr[0, :4])
