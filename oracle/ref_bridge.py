"""Import the UNMODIFIED reference (jinPrelude/simple-es) from /root/reference.

TEST INFRASTRUCTURE ONLY; usable only in the dev container (the GPU box has no
/root/reference).  Used by oracle/make_golden.py to generate tests/golden/*.npz and by
dev-container-only tests that compare the ports in oracle/pyref.py with the real classes.

The reference's env wrappers need gym / pybullet / pettingzoo, none of which is installed, so
only the modules that import cleanly are exposed; environments come from oracle/pyref.py shims
that satisfy the wrapper duck type.
"""
import contextlib
import os
import sys

REF_ROOT = os.environ.get("SES_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "learning_strategies"))


def load():
    """Returns a namespace with the reference's strategies, Adam, GymEnvModel, RolloutWorker, ESLoop."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    sys.dont_write_bytecode = True
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    os.environ.setdefault("WANDB_MODE", "disabled")
    from learning_strategies.evolution import offspring_strategies as strategies
    from learning_strategies.evolution import loop
    from learning_strategies.evolution import utils as es_utils
    from learning_strategies import optimizers
    from networks import neural_network

    class NS:
        pass

    ns = NS()
    ns.strategies = strategies
    ns.simple_evolution = strategies.simple_evolution
    ns.simple_genetic = strategies.simple_genetic
    ns.openai_es = strategies.openai_es
    ns.Adam = optimizers.Adam
    ns.GymEnvModel = neural_network.GymEnvModel
    ns.RolloutWorker = loop.RolloutWorker
    ns.ESLoop = loop.ESLoop
    ns.wrap_agentid = es_utils.wrap_agentid
    return ns


@contextlib.contextmanager
def stable_argsort(strategies_module):
    """Pin np.argsort's tie order to kind='stable' while the reference's evaluate() runs
    (the default introsort/SIMD sort tie order is build dependent)."""
    np_mod = strategies_module.np
    orig = np_mod.argsort

    def pinned(a, *args, **kw):
        kw.setdefault("kind", "stable")
        return orig(a, *args, **kw)

    np_mod.argsort = pinned
    try:
        yield
    finally:
        np_mod.argsort = orig


@contextlib.contextmanager
def capture_locals(func_name, sink):
    """Record the local variables of the reference function `func_name` at its return."""
    def prof(frame, event, arg):
        if event == "return" and frame.f_code.co_name == func_name and REF_ROOT in frame.f_code.co_filename:
            sink.append(dict(frame.f_locals))
    old = sys.getprofile()
    sys.setprofile(prof)
    try:
        yield
    finally:
        sys.setprofile(old)
