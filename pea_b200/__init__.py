"""Importable alias of the ``pixel-embedded-affinity_b200`` directory (a hyphen cannot appear in an import).

``import pea_b200`` executes that directory's ``__init__.py`` with this package's ``__path__`` pointed at it, so
``pea_b200.loss_embedding_mse_2d`` etc. resolve to the files there.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "pixel-embedded-affinity_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
