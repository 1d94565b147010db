"""build_loop with the engine switch (the reference's seam: builder.py:27-86).

``engine: {name: b200}`` in the YAML selects the GPU loop; without it the call is handed to the
reference's own builder, unchanged, if the reference checkout is importable (SES_REFERENCE_ROOT or
sys.path).  There is deliberately no CPU re-implementation behind this function.
"""
import importlib
import os
import sys

from .loop import B200Loop


def engine_name(config):
    eng = config.get("engine") or {}
    return eng.get("name") if isinstance(eng, dict) else eng


def build_loop(config, gen_num, process_num, eval_ep_num, log, save_model_period, seed=0):
    if engine_name(config) == "b200":
        return B200Loop(config, gen_num, process_num, eval_ep_num, log, save_model_period, seed=seed)
    root = os.environ.get("SES_REFERENCE_ROOT")
    if root and root not in sys.path:
        sys.path.insert(0, root)
    try:
        ref_builder = importlib.import_module("builder")
    except Exception as exc:       # gym / pybullet / pettingzoo missing, or no reference checkout
        raise RuntimeError(
            "config has no `engine: {name: b200}` key and the reference builder is not importable "
            "(set SES_REFERENCE_ROOT to a simple-es checkout with its dependencies): %s" % (exc,))
    if getattr(ref_builder, "build_loop", None) is build_loop or not hasattr(ref_builder, "build_env"):
        raise RuntimeError("config has no `engine: {name: b200}` key and no reference builder is on sys.path")
    return ref_builder.build_loop(config, gen_num, process_num, eval_ep_num, log, save_model_period)
