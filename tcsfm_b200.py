"""Importable alias for the package directory ``tightly-coupled-sfm_b200/``.

The directory name is fixed by the build contract and contains hyphens, so it
cannot be the target of an ``import`` statement.  This module turns itself into
that package: ``import tcsfm_b200`` / ``from tcsfm_b200 import stn, losses``.
"""
import os as _os

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "tightly-coupled-sfm_b200")
__path__ = [_pkg_dir]
__file__ = _os.path.join(_pkg_dir, "__init__.py")
with open(__file__, "r") as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _f
