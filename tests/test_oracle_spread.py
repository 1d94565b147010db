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


def test_spread_contact_forces_conserve_momentum_and_repel(twin):
    """Independent physical pins of the restated MPE world (pettingzoo is not installable here): contact forces are equal and
    opposite, so with every agent idle (all-zero policy -> action 0) the total momentum stays zero although overlapping
    agents are pushed apart; the penetration force is repulsive; and the episode return equals the reward formula of
    SURVEY.md Appendix A.2 re-evaluated in numpy from the traced positions."""
    N, D = 3, 773
    init = np.array([0.10, 0.00, 0.22, 0.05, -0.60, -0.55,          # agents 0 and 1 start overlapping (dist 0.13 < 0.3), agent 2 far
                     0.5, 0.5, -0.5, 0.5, 0.0, -0.8], dtype=np.float64)[None]  # landmarks
    f, n, tr, ac = twin.rollout_mpe(np.zeros(D, np.float32), N=N, E=1, init=init, trace_steps=25)
    assert n == 25 and np.all(ac == 0)
    pos, vel = tr[:, :2 * N].reshape(25, N, 2), tr[:, 2 * N:].reshape(25, N, 2)
    assert np.abs(vel.sum(axis=1)).max() < 1e-12                      # sum_i m v_i == 0 (masses are 1)
    d01 = np.linalg.norm(pos[:, 0] - pos[:, 1], axis=1)
    assert d01[0] > 0.13 and np.all(np.diff(d01) > 0)                 # the overlapping pair separates monotonically
    assert np.abs(vel[0, 2]).max() < 1e-3                             # the distant agent feels (almost) nothing
    lm = init[0, 2 * N:].reshape(N, 2)
    total = 0.0
    for t in range(25):
        glob = -sum(min(np.linalg.norm(pos[t, a] - lm[l]) for a in range(N)) for l in range(N))
        for i in range(N):
            local = -sum(1.0 for a in range(N) if np.linalg.norm(pos[t, a] - pos[t, i]) < 0.3)    # counts the agent itself
            total += 0.5 * glob + 0.5 * local
    assert abs(f - total) <= 1e-12 * abs(total)
