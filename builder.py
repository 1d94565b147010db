"""Top-level ``builder`` module, as in the reference layout (builder.py:27): ``build_loop`` with the
engine switch.  The implementation lives in simple-es_b200/builder.py."""
from simple_es_b200.builder import build_loop, engine_name  # noqa: F401
