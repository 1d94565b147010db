"""N > 1 on real GPUs: population sharded over 2 ranks (NCCL) must reproduce the 1-GPU run bit for
bit -- fitness, rank order and the updated parameters on EVERY rank (shared Philox seeds, no
parameter traffic).  Needs >= 2 visible GPUs (run with `gpurun --gpus 2`); skipped otherwise."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, strategy, exchange, q, shard="cyclic"):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    if exchange == "peer_fold":      # test build: the flag barriers inside K1 (sentinel CTA) and k_grad_partial (last CTA)
        os.environ.update(SES_B200_TEST_BUILD="1", SES_PEER_FOLD="1")
        exchange = "peer"
    import yaml
    from simple_es_b200.loop import B200Loop
    cfg = yaml.load(open(os.path.join(ROOT, "conf", {"openai_es": "cartpole_openai.yaml", "simple_evolution": "cartpole.yaml"}[strategy])),
                    Loader=yaml.FullLoader)
    cfg["strategy"]["offspring_num"] = 6000 if strategy == "openai_es" else 3000     # evolution: P = 3001 (ragged shards)
    cfg["engine"]["fitness_exchange"] = exchange
    cfg["engine"]["shard"] = shard
    loop = B200Loop(cfg, 4, 1, 5, save_model_period=0, seed=3, quiet=True)
    for _ in range(4):
        loop.strategy.step()
    s = loop.strategy
    torch.cuda.synchronize()
    q.put((rank, s.parents.cpu().numpy(), s.fitness.cpu().numpy(), s.order.cpu().numpy(), int(s.total_env_steps.item())))
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


def _single(strategy):
    import yaml
    from simple_es_b200.loop import B200Loop
    cfg = yaml.load(open(os.path.join(ROOT, "conf", {"openai_es": "cartpole_openai.yaml", "simple_evolution": "cartpole.yaml"}[strategy])),
                    Loader=yaml.FullLoader)
    cfg["strategy"]["offspring_num"] = 6000 if strategy == "openai_es" else 3000
    loop = B200Loop(cfg, 4, 1, 5, save_model_period=0, seed=3, quiet=True)
    for _ in range(4):
        loop.strategy.step()
    s = loop.strategy
    return s.parents.cpu().numpy(), s.fitness.cpu().numpy(), s.order.cpu().numpy(), int(s.total_env_steps.item())


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("strategy,exchange,shard", [("openai_es", "nccl", "cyclic"), ("simple_evolution", "nccl", "cyclic"),
                                                     ("openai_es", "peer", "cyclic"), ("simple_evolution", "peer", "cyclic"),
                                                     ("openai_es", "nccl", "contiguous"), ("simple_evolution", "peer", "contiguous"),
                                                     ("openai_es", "peer_fold", "cyclic")])
def test_two_rank_run_equals_single_gpu(strategy, exchange, shard):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 200) + (7 if exchange == "peer" else 3 if exchange == "peer_fold" else 0) + (13 if strategy == "openai_es" else 0) + (29 if shard == "cyclic" else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, strategy, exchange, q, shard)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    par1, fit1, ord1, steps1 = _single(strategy)
    for rank, par, fit, order, steps in res:
        assert np.array_equal(fit, fit1) and np.array_equal(order, ord1)      # identical fitness / selection on every rank
        assert np.array_equal(par, par1)                                      # identical update, nothing was broadcast
    assert res[0][4] + res[1][4] == steps1                                    # every env step simulated exactly once
