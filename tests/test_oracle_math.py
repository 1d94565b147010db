"""The oracle's elementary functions and RNG against independent references (numpy/libm/KAT)."""
import math

import numpy as np


def test_philox_known_answers(twin):
    # Random123 kat_vectors, philox4x32-10
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, want in kat:
        assert [int(v) for v in twin.philox(ctr, key)] == want


def test_tanh_accuracy(twin):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-9.5, 9.5, 1_000_000), rng.normal(0, 1, 1_000_000),
                        10.0 ** rng.uniform(-6, 0, 200_000)]).astype(np.float32)
    t = np.tanh(x.astype(np.float64))
    y = twin.tanhf(x).astype(np.float64)
    ulp = np.spacing(np.abs(t).astype(np.float32)).astype(np.float64)
    assert (np.abs(y - t) / ulp).max() < 6.0
    assert np.abs(y - t).max() < 4e-7
    assert twin.tanhf([0.0])[0] == 0.0
    big = twin.tanhf([50.0, -50.0, np.inf, -np.inf])
    assert np.all(np.abs(np.abs(big) - 1.0) < 2e-7)
    s = twin.sigmf(x).astype(np.float64)
    assert np.abs(s - 1.0 / (1.0 + np.exp(-x.astype(np.float64)))).max() < 3e-7


def test_ln_sincos2pi(twin):
    rng = np.random.default_rng(1)
    u = ((rng.integers(0, 2 ** 24, 500_000) + 0.5) * 2.0 ** -24).astype(np.float32)
    l = twin.lnf(u).astype(np.float64)
    assert np.abs(l - np.log(u.astype(np.float64))).max() < 2e-6
    assert twin.lnf([1.0])[0] == 0.0
    v = (rng.integers(0, 2 ** 24, 500_000) * 2.0 ** -24).astype(np.float32)
    s, c = twin.sincos2pif(v)
    assert np.abs(s - np.sin(2 * np.pi * v.astype(np.float64))).max() < 2e-7
    assert np.abs(c - np.cos(2 * np.pi * v.astype(np.float64))).max() < 2e-7


def test_sincos_f64_within_one_ulp_of_libm(twin):
    rng = np.random.default_rng(2)
    th = np.concatenate([rng.uniform(-0.5, 0.5, 100_000), rng.uniform(-0.21, 0.21, 100_000), [0.0]])
    s, c = twin.sincos(th)
    ms = np.array([math.sin(a) for a in th])
    mc = np.array([math.cos(a) for a in th])
    assert np.all(np.abs(s - ms) <= np.spacing(np.abs(ms)))
    assert np.all(np.abs(c - mc) <= np.spacing(np.abs(mc)))
    assert (s != ms).mean() < 0.02 and (c != mc).mean() < 0.03


def test_normals_distribution(twin):
    from scipy import stats
    n = np.stack([twin.normals(7, 3, i, 226) for i in range(3000)])
    flat = n.ravel().astype(np.float64)
    assert abs(flat.mean()) < 5e-3 and abs(flat.std() - 1) < 5e-3
    assert abs(((flat - flat.mean()) ** 4).mean() / flat.var() ** 2 - 3.0) < 0.03
    assert stats.kstest(flat, "norm").pvalue > 1e-3
    # independent across offspring / parameters / generations / seeds
    assert abs(np.corrcoef(n[:-1].ravel(), n[1:].ravel())[0, 1]) < 5e-3
    assert abs(np.corrcoef(n[:, :-1].ravel(), n[:, 1:].ravel())[0, 1]) < 5e-3
    assert not np.array_equal(twin.normals(7, 3, 0, 226), twin.normals(7, 4, 0, 226))
    assert not np.array_equal(twin.normals(7, 3, 0, 226), twin.normals(8, 3, 0, 226))
    # counter based: a prefix of a longer vector is the shorter vector
    assert np.array_equal(twin.normals(7, 3, 5, 226), twin.normals(7, 3, 5, 6562)[:226])
