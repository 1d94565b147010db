"""build_loop with the engine switch (the reference's seam: builder.py:27-86).

``engine: {name: b200}`` in the YAML selects the GPU loop; without it the call is handed to the
reference's own builder, unchanged, if the reference checkout is importable (SES_REFERENCE_ROOT or
sys.path).  There is deliberately no CPU re-implementation behind this function.
"""
import importlib
import importlib.util
import os
import sys

from .loop import B200Loop


def engine_name(config):
    eng = config.get("engine") or {}
    return eng.get("name") if isinstance(eng, dict) else eng


def build_loop(config, gen_num, process_num, eval_ep_num, log, save_model_period, seed=0):
    if engine_name(config) == "b200":
        return B200Loop(config, gen_num, process_num, eval_ep_num, log, save_model_period, seed=seed)
    return _reference_builder().build_loop(config, gen_num, process_num, eval_ep_num, log, save_model_period)


_REF_BUILDER = None


def _reference_builder():
    """The reference checkout's own builder.py, loaded BY PATH: `import builder` would return this repository's
    top-level shim (run_es.py / sweep_main.py import it under that very name), never the reference's module.  The
    checkout's root goes first on sys.path while the module executes so that its `envs`, `networks` and
    `learning_strategies` imports resolve there."""
    global _REF_BUILDER
    if _REF_BUILDER is not None:
        return _REF_BUILDER
    root = os.environ.get("SES_REFERENCE_ROOT")
    if not root:
        raise RuntimeError("config has no `engine: {name: b200}` key: set SES_REFERENCE_ROOT to a simple-es checkout (with its "
                           "dependencies installed) to run it on the reference CPU path")
    path = os.path.join(root, "builder.py")
    if not os.path.isfile(path):
        raise RuntimeError("SES_REFERENCE_ROOT=%s holds no builder.py" % (root,))
    spec = importlib.util.spec_from_file_location("ses_reference_builder", path)
    mod = importlib.util.module_from_spec(spec)
    sys.path.insert(0, root)
    try:
        spec.loader.exec_module(mod)
    except Exception as exc:       # gym / pybullet / pettingzoo missing
        raise RuntimeError("config has no `engine: {name: b200}` key and the reference builder at %s cannot be imported "
                           "(are the reference's dependencies installed?): %s: %s" % (path, type(exc).__name__, exc))
    finally:
        if sys.path and sys.path[0] == root:
            sys.path.pop(0)
        if root not in sys.path:
            sys.path.append(root)       # the reference's lazy imports (its modules import each other by top-level name)
    if not hasattr(mod, "build_loop"):
        raise RuntimeError("%s has no build_loop" % (path,))
    _REF_BUILDER = mod
    return mod
