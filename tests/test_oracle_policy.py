"""The bit-twin policy against the reference's GymEnvModel.forward (torch CPU) outputs."""
import numpy as np


def _check(twin, g, tol):
    W, O, A, Z = g["W"], g["obs"], g["actions"], g["logits"]
    gru = bool(g["gru"])
    mism, worst = 0, 0.0
    for i in range(W.shape[0]):
        h = None
        for t in range(O.shape[1]):
            a, z, h = twin.policy_step(W[i], 4, 2, gru, O[i, t].astype(np.float32), h)
            mism += int(a != A[i, t])
            worst = max(worst, float(np.abs(z - Z[i, t]).max() / max(1.0, np.abs(Z[i, t]).max())))
    assert mism <= 1e-3 * A.size
    assert worst < tol


def test_policy_mlp_golden(twin, golden):
    _check(twin, golden("policy_mlp"), 1e-5)


def test_policy_gru_golden(twin, golden):
    _check(twin, golden("policy_gru"), 1e-4)


def test_softmax_collapse_rule(twin):
    """argmax(softmax(z)) picks the lowest index whenever exp(z_j - z_max) rounds to 1.0f."""
    import torch
    w = np.zeros(226, np.float32)
    cases = [(0.0, 0.0), (0.0, 2e-8), (0.0, 4e-8), (1e-3, 1e-3 + 2e-8), (5.0, 5.0 + 1e-6), (-3.0, -3.0 + 2.5e-7), (2.0, 1.0)]
    for z0, z1 in cases:
        w[224], w[225] = z0, z1
        a, z, _ = twin.policy_step(w, 4, 2, False, np.zeros(4, np.float32))
        want = int(torch.argmax(torch.softmax(torch.tensor([np.float32(z0), np.float32(z1)]), dim=0)))
        assert a == want, (z0, z1, a, want)
