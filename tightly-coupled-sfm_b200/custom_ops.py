"""`torch.library` registrations of the two tensor-in / tensor-out operators of the path, so that callers who
trace (fake tensors, `torch.compile`, `torch.export`) see opaque, shape-inferable operators instead of a
`ctypes` call:

    torch.ops.tcsfm.inverse_warp2(img, depth, ref_depth, kinv, proj) -> (projected_img, valid_mask,
                                                                         projected_depth, computed_depth)
    torch.ops.tcsfm.ssim(x, y) -> dissimilarity map

Each is a `torch.library.custom_op` over the C ABI (include/tcsfm.h) with a fake (meta) implementation and a
hand-written backward registered through `register_autograd` (the backward is itself a custom op, so double
tracing works).  The eager drop-ins (`stn.inverse_warp2`, `losses.SSIM_Loss`) keep using the thin
`autograd.Function`s of ops.py, which additionally skip unused outputs / gradients; `stn.inverse_warp2_op` and
`losses.ssim_op` are the traceable spellings.  No CPU fallback: the implementations call the CUDA library.
"""
from typing import Tuple

import torch

from . import _raw, ops


def _lib():
    return ops.lib()


@torch.library.custom_op("tcsfm::inverse_warp2", mutates_args=())
def inverse_warp2_op(img: torch.Tensor, depth: torch.Tensor, ref_depth: torch.Tensor, kinv: torch.Tensor,
                     proj: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """models/stn.py:265-273 given K^-1 [B,3,3] and K[R|t] [B,3,4]."""
    ops._require_cuda(img, depth, ref_depth, kinv, proj)
    flags = ops.arith_flags(img.shape[0], img.shape[2], img.shape[3])
    with ops._guard(img):
        return _raw.warp_fwd(_lib(), img, depth, ref_depth, kinv, proj, flags)


@inverse_warp2_op.register_fake
def _(img, depth, ref_depth, kinv, proj):
    b, _, h, w = img.shape
    one = img.new_empty((b, 1, h, w))
    return img.new_empty((b, 3, h, w)), one, torch.empty_like(one), torch.empty_like(one)


@torch.library.custom_op("tcsfm::inverse_warp2_backward", mutates_args=())
def inverse_warp2_backward_op(img: torch.Tensor, depth: torch.Tensor, ref_depth: torch.Tensor, kinv: torch.Tensor,
                              proj: torch.Tensor, g_img: torch.Tensor, g_pd: torch.Tensor,
                              g_cd: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    flags = ops.arith_flags(img.shape[0], img.shape[2], img.shape[3])
    with ops._guard(img):
        g_depth, g_ref, g_proj, _ = _raw.warp_bwd(_lib(), img, depth, ref_depth, kinv, proj, g_img, g_pd, g_cd, flags)
    return g_depth, g_ref, g_proj


@inverse_warp2_backward_op.register_fake
def _(img, depth, ref_depth, kinv, proj, g_img, g_pd, g_cd):
    return torch.empty_like(depth), torch.empty_like(ref_depth), torch.empty_like(proj)


def _warp_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs)


def _warp_backward(ctx, g_img, g_valid, g_pd, g_cd):
    img, depth, ref_depth, kinv, proj = ctx.saved_tensors
    g_depth, g_ref, g_proj = inverse_warp2_backward_op(img, depth, ref_depth, kinv, proj, g_img.contiguous(),
                                                       g_pd.contiguous(), g_cd.contiguous())
    return None, g_depth, g_ref, None, g_proj


inverse_warp2_op.register_autograd(_warp_backward, setup_context=_warp_setup)


@torch.library.custom_op("tcsfm::ssim", mutates_args=())
def ssim_op(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """SSIM_Loss.forward (losses.py:27-41)."""
    ops._require_cuda(x, y)
    with ops._guard(x):
        return _raw.ssim_fwd(_lib(), x, y, ops.ARITH_FLAGS)


@ssim_op.register_fake
def _(x, y):
    return torch.empty_like(x)


@torch.library.custom_op("tcsfm::ssim_backward", mutates_args=())
def ssim_backward_op(x: torch.Tensor, y: torch.Tensor, g_out: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    with ops._guard(x):
        return _raw.ssim_bwd(_lib(), x, y, g_out, True, True, ops.ARITH_FLAGS)


@ssim_backward_op.register_fake
def _(x, y, g_out):
    return torch.empty_like(x), torch.empty_like(y)


def _ssim_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs)


def _ssim_backward(ctx, g_out):
    x, y = ctx.saved_tensors
    return ssim_backward_op(x, y, g_out.contiguous())


ssim_op.register_autograd(_ssim_backward, setup_context=_ssim_setup)
