"""Importable alias for the product package, which lives in ``llava-reward_b200/``.

The directory name the build contract asks for (``llava-reward_b200``) is not a
valid Python identifier, so this shim points the ``llava_reward_b200`` package
path at that directory and executes its ``__init__``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "llava-reward_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
