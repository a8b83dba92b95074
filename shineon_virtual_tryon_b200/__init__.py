"""Importable alias of the ``shineon-virtual-tryon_b200/`` package directory.

The package directory carries the repository's name (with hyphens), which Python cannot import
directly; this stub makes ``import shineon_virtual_tryon_b200`` resolve to it.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "shineon-virtual-tryon_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
