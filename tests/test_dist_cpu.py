"""Host-side sharding logic on CPU: world_size-2 gloo processes (the N > 1 path without a GPU)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, P, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from simple_es_b200 import dist as sdist
    from simple_es_b200.engine import shard_bounds
    r, w = sdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    lo, hi = shard_bounds(P, rank, world)
    full = np.random.RandomState(0).uniform(-3, 500, P)           # what an unsharded run would produce
    fit = torch.full((P,), float("nan"), dtype=torch.float64)      # stale garbage outside the slice
    fit[lo:hi] = torch.from_numpy(full[lo:hi])
    sdist.exchange_fitness(fit, lo, hi)
    total = sdist.sum_scalar(torch.tensor(hi - lo, dtype=torch.int64))
    q.put((rank, lo, hi, bool(np.array_equal(fit.numpy(), full)), int(total)))
    dist.destroy_process_group()


@pytest.mark.parametrize("P", [65536, 97, 4097])
def test_fitness_exchange_world2_gloo(P):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() + P) % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, P, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, ok0, t0), (r1, lo1, hi1, ok1, t1) = res
    assert lo0 == 0 and hi0 == lo1 and hi1 == P and abs((hi0 - lo0) - (hi1 - lo1)) <= 1    # contiguous, balanced
    assert ok0 and ok1                                                                      # bit-identical on every rank
    assert t0 == t1 == P


def _worker_cyclic(rank, world, port, P, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from simple_es_b200 import dist as sdist
    from simple_es_b200.engine import cyclic_block, owned_ids
    sdist.init_from_env(backend="gloo")
    mine = owned_ids(P, rank, world, cyclic_block(P, world))
    full = np.random.RandomState(1).uniform(-300, 500, P)
    fit = torch.full((P,), float("nan"), dtype=torch.float64)
    fit[torch.from_numpy(mine)] = torch.from_numpy(full[mine])
    mask = torch.ones(P, dtype=torch.bool); mask[torch.from_numpy(mine)] = False
    sdist.exchange_fitness_masked(fit, mask)
    q.put((rank, int(mine.size), bool(np.array_equal(fit.numpy(), full))))
    dist.destroy_process_group()


@pytest.mark.parametrize("P", [65536, 97, 4097])
def test_fitness_exchange_block_cyclic_world2_gloo(P):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29950 + (os.getpid() + P) % 40
    procs = [ctx.Process(target=_worker_cyclic, args=(r, 2, port, P, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] + res[1][1] == P and res[0][2] and res[1][2]


def test_block_cyclic_ownership_partitions_the_population():
    sys.path.insert(0, ROOT)
    from simple_es_b200.engine import cyclic_block, owned_ids
    for P in (2, 97, 4097, 65536, (1 << 20) + 5):
        for world in (1, 2, 3, 8):
            B = cyclic_block(P, world)
            parts = [owned_ids(P, r, world, B) for r in range(world)]
            allids = np.sort(np.concatenate(parts))
            assert np.array_equal(allids, np.arange(P))                                   # a partition
            if P >= 64 * world:
                assert max(p.size for p in parts) - min(p.size for p in parts) <= B        # balanced to one block
            for r, ids in enumerate(parts):                                               # the device-side map (Shard::local_to_id)
                l = np.arange(ids.size)
                assert np.array_equal(((l // B) * world + r) * B + l % B, ids)
                assert np.array_equal(((ids // B) // world) * B + ids % B, l)             # ... and its inverse


def test_shard_bounds_cover_population():
    sys.path.insert(0, ROOT)
    from simple_es_b200.engine import shard_bounds, population_layout
    for P in (2, 97, 4097, 65536, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            b = [shard_bounds(P, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == P and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
    # population sizes of the three strategies
    assert population_layout("simple_evolution", 96, 10) == (97, 97, 2, 1)
    assert population_layout("openai_es", 65536) == (65536, 65536, 1, 1)
    assert population_layout("simple_genetic", 1048576, 16) == (1048576, 65536, 1, 16)
    assert population_layout("simple_genetic", 26, 4) == (24, 6, 1, 4)
