#!/usr/bin/env python
"""wandb-sweep entry point with the reference's command line (sweep_main.py:33-91 of jinPrelude/simple-es).

    wandb sweep sweep_config/cartpole_openaies.yaml        # prints:  wandb agent <sweep id>
    wandb agent <sweep id>                                 # runs:    python sweep_main.py --cfg-path=... --init-sigma=... ...

The sweep agent passes hyper-parameters as flags; every flag that is given replaces the value of the YAML key of
the same name (dashes -> underscores), wherever that key sits in the config -- the behaviour of the reference's
``change_value`` (sweep_main.py:16-30).  The resulting config goes through the same ``builder.build_loop`` seam as
run_es.py, so ``engine: {name: b200}`` in the YAML makes every sweep trial run on the GPU engine.
"""
import argparse

import yaml

import builder
import run_es

# hyper-parameters a sweep may override (sweep_main.py:65-69): flag -> type
OVERRIDES = (("--init-sigma", float), ("--sigma-decay", float), ("--learning-rate", float), ("--elite-num", int),
             ("--offspring-num", int))


def parse_args(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    for name, typ, default, text in run_es.FLAGS:
        if name == "--generation-num":
            default = 1000                      # sweep_main.py:50-55
        if name == "--cfg-path":
            default = "conf/cartpole_openai.yaml"
        ap.add_argument(name, type=typ, default=default, help=text)
    # the reference logs to wandb unless --log is given (store_false, sweep_main.py:62): a sweep needs the metric
    ap.add_argument("--log", action="store_false", help="disable wandb logging")
    for name, typ in OVERRIDES:
        ap.add_argument(name, type=typ, default=None)
    return ap.parse_args(argv)


def apply_overrides(config, overrides):
    """Set config[section]...[key] = value for every (key, value) with value not None, in every section that
    already has `key` (a sweep cannot invent keys: a strategy without `elite_num` ignores --elite-num).  The search
    descends into nested dicts; returns the list of dotted paths that were changed."""
    changed = []

    def visit(node, path):
        for k, v in node.items():
            if isinstance(v, dict):
                visit(v, path + [k])
        for key, value in overrides.items():
            if value is not None and key in node and not isinstance(node[key], dict):
                node[key] = value
                changed.append(".".join(path + [key]))

    visit(config, [])
    return changed


def main(argv=None):
    args = parse_args(argv)
    with open(args.cfg_path) as fh:
        config = yaml.load(fh, Loader=yaml.FullLoader)
    overrides = {name.lstrip("-").replace("-", "_"): getattr(args, name.lstrip("-").replace("-", "_")) for name, _ in OVERRIDES}
    apply_overrides(config, overrides)
    run_es.seed_everything(args.seed)
    loop = builder.build_loop(config, args.generation_num, args.process_num, args.eval_ep_num, args.log,
                              args.save_model_period, seed=args.seed)
    loop.run()
    return loop


if __name__ == "__main__":
    main()
