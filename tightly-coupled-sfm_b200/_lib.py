"""Loader of the CUDA C-ABI library.  There is deliberately no fallback: if
libtcsfm_b200.so is missing the package raises, it never computes on the CPU."""
import os

from . import _cabi

_LIB = None
# TCSFM_B200_LIB points at an alternative build of the same CUDA library (kernel tuning
# experiments); it is still a CUDA build of csrc/ -- there is no non-CUDA implementation.
LIB_PATH = os.environ.get("TCSFM_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                                             "libtcsfm_b200.so")


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                "tcsfm_b200: %s not found. Build it with `python -m tcsfm_b200.build` (nvcc, sm_100a); "
                "there is no CPU fallback." % LIB_PATH)
        _LIB = _cabi.bind(LIB_PATH)
    return _LIB
