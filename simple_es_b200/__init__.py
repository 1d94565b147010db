"""simple-es B200 population-rollout engine (import name of the ``simple-es_b200/`` product directory).

Drop-in for the rollout hot path of jinPrelude/simple-es (perturb -> rollout -> fitness -> rank/select -> update) as
hand-written sm_100a CUDA kernels behind the C ABI in ``include/ses_b200.h``.  There is no CPU fallback: importing works
anywhere, but every compute call needs ``libses_b200.so`` and a CUDA device and fails loudly otherwise.

The product directory carries the project's name, whose hyphen Python cannot import; this is a regular package whose
search path is extended with that directory (the mechanism of ``pkgutil.extend_path``), so ``simple_es_b200.engine`` etc.
are ordinary submodules found by the import system -- no code is executed from here.
"""
import os as _os

__version__ = "0.2.0"
__path__.append(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "simple-es_b200"))
