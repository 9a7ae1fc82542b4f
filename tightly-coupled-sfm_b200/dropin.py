"""Installs the drop-ins over the reference's module names so that an unmodified
reference checkout (``train_mono.py``, ``optimization_experiments/*.py``) calls the
fused kernels.  Call before importing the reference's scripts."""
import sys
import types


def install():
    from . import geometry_helpers, losses, stn
    models_pkg = sys.modules.get("models")
    if models_pkg is None:
        try:
            import models as models_pkg          # the reference's package, if it is on sys.path
        except ImportError:
            models_pkg = types.ModuleType("models")
            models_pkg.__path__ = []
            sys.modules["models"] = models_pkg
    models_pkg.stn = stn
    sys.modules["models.stn"] = stn
    utils_pkg = sys.modules.get("utils")
    if utils_pkg is None:
        try:
            import utils as utils_pkg
        except ImportError:
            utils_pkg = types.ModuleType("utils")
            utils_pkg.__path__ = []
            sys.modules["utils"] = utils_pkg
    utils_pkg.geometry_helpers = geometry_helpers
    sys.modules["utils.geometry_helpers"] = geometry_helpers
    sys.modules["losses"] = losses
    return {"models.stn": stn, "utils.geometry_helpers": geometry_helpers, "losses": losses}
