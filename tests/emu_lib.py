"""TEST INFRASTRUCTURE ONLY: binds the g++-compiled CUDA-emulator build of the
kernels (tests/emu).  Lets `-m "not gpu"` tests exercise kernel logic in the
GPU-less container.  Never imported by the package."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))

_LIB = None


def emu():
    global _LIB
    if _LIB is None:
        import build_emu
        from tcsfm_b200 import _cabi
        _LIB = _cabi.bind(build_emu.build())
    return _LIB
