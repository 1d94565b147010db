#!/usr/bin/env python
"""Training entry point with the reference's command line (run_es.py:15-62 of jinPrelude/simple-es).

    python run_es.py --cfg-path conf/cartpole.yaml --generation-num 100
    torchrun --nproc-per-node 8 run_es.py --cfg-path conf/cartpole_openai.yaml      # population sharded over 8 GPUs

Same flags, same YAML schema; a config with ``engine: {name: b200}`` runs on the GPU engine, any other
config is handed to the reference's own builder when a reference checkout is importable.
"""
import argparse
import random

import numpy as np
import torch
import yaml

import builder

FLAGS = (
    # name, type, default, help  -- the reference's flags, unchanged (run_es.py:17-45)
    ("--cfg-path", str, "conf/cartpole.yaml", "config file to run."),
    ("--seed", int, 0, "random seed."),
    ("--process-num", int, 12, "number of mp process (ignored by the GPU engine)."),
    ("--generation-num", int, 10000, "max number of generation iteration."),
    ("--eval-ep-num", int, 5, "number of model evaluaion per iteration."),
    ("--save-model-period", int, 10, "save model for every n iteration."),
)


def parse_args(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    for name, typ, default, text in FLAGS:
        ap.add_argument(name, type=typ, default=default, help=text)
    ap.add_argument("--log", action="store_true", help="wandb log")
    return ap.parse_args(argv)


def seed_everything(seed):
    for fn in (torch.manual_seed, np.random.seed, random.seed):
        fn(seed)


def main(argv=None):
    args = parse_args(argv)
    seed_everything(args.seed)
    with open(args.cfg_path) as fh:
        config = yaml.load(fh, Loader=yaml.FullLoader)
    loop = builder.build_loop(config, args.generation_num, args.process_num, args.eval_ep_num, args.log,
                              args.save_model_period, seed=args.seed)
    loop.run()
    return loop


if __name__ == "__main__":
    main()
