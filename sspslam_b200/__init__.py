"""Importable alias for the package directory ``semantic-spiking-neural-slam-2023_b200/``
(a hyphenated directory name cannot be imported directly)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "semantic-spiking-neural-slam-2023_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _f
