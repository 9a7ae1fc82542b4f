"""tcsfm-b200: fused sm_100a inverse-warp + SSIM/L1 photometric loss.

Drop-in replacements for the reference's ``models/stn.py``, ``losses.py`` and
``utils/geometry_helpers.py`` call signatures, backed by hand-written CUDA
kernels behind a C ABI (``include/tcsfm.h``).  There is no CPU fallback: every
operator raises if the CUDA library is missing or the tensors are not on a GPU.
"""
__version__ = "0.1.0"
