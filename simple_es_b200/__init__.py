"""Importable alias of the ``simple-es_b200/`` package directory.

The product directory carries the name the project was given (with a hyphen, which Python cannot
import); this stub makes ``import simple_es_b200`` resolve every submodule from there.
"""
import os as _os

_impl = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "simple-es_b200")
__path__ = [_impl]
with open(_os.path.join(_impl, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_impl, "__init__.py"), "exec"))
