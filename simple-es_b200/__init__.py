"""simple-es B200 population-rollout engine.

Drop-in for the rollout hot path of jinPrelude/simple-es (perturb -> rollout -> fitness ->
rank/select -> update) as hand-written sm_100a CUDA kernels behind the C ABI in
``include/ses_b200.h``.  There is no CPU fallback: importing works anywhere, but every compute
call needs ``libses_b200.so`` and a CUDA device and fails loudly otherwise.
"""
__version__ = "0.1.0"
