"""Drop-ins for the loss half of the reference's per-frame test-time optimisation
(PFT): ``DepthOptimizer.compute_optimization_loss`` (optimization_experiments/
optimizer.py:29-97) and ``compute_photometric_error`` (optimization_experiments/
helpers.py:8-23).  The optimisation driver itself (deep copies of the depth net,
Adam, plotting; optimizer.py:136-297) is the reference's harness and calls these.
"""
import torch

from . import _cabi, ops
from .losses import get_smooth_loss
from .stn import inverse_warp2


def _stacked(fwd_data, inv_data, key):
    """The [2*S*B,1,H,W] tensor whose two halves are fwd_data[key] / inv_data[key].  The dictionaries
    solve_pose_iteratively returns hold slices of one stacked tensor (train_mono.py:94-100): then the base is
    used as it is (no copy, and the gradient flows straight into it); anything else is concatenated once."""
    f, i = fwd_data[key], inv_data[key]
    base = f._base
    if (base is not None and base is i._base and base.is_contiguous() and f.shape == i.shape
            and base.shape[0] == 2 * f.shape[0] and base.shape[1:] == f.shape[1:]
            and f.data_ptr() == base.data_ptr() and i.data_ptr() == base.data_ptr() + f.numel() * f.element_size()):
        return base
    return torch.cat([f, i], 0)


def _pft_flags(options):
    flags = 0
    if options['diff_img_argmin'] == True:             # noqa: E712 -- the reference compares with ==
        flags |= _cabi.PFT_ARGMIN
    if options['automasking'] == True:                 # noqa: E712
        flags |= _cabi.PFT_AUTOMASK
    if options['l_inverse_reconstruction'] == True:    # noqa: E712
        flags |= _cabi.PFT_INVERSE
    if options['l_depth_consist'] == True:             # noqa: E712
        flags |= _cabi.PFT_DEPTH_CONSIST
    return flags


def compute_optimization_loss(options, target_img, target_disparity, init_disparity, fwd_data, inv_data,
                              ssim_loss=None):
    """optimizer.py:45-97 without the plotting branches.  `init_disparity` is the reference's
    ``self.target_disparity`` (the un-optimised prediction).

    The reconstruction and depth-consistency terms (optimizer.py:45-86: per-pixel min over the sources,
    valid-mask union, auto-mask, the masked weighted means and the two (1 - weight).mean() terms) are one
    fused launch forward and one backward (csrc/pft_kernels.cu); the depth-initialisation term is the fused
    SSIM-mean pair; no full-image PyTorch operator runs here."""
    bsz = target_img.shape[0]
    n_src = options['num_source_imgs']
    if fwd_data['diff_img'].shape[0] != n_src * bsz or inv_data['diff_img'].shape[0] != n_src * bsz:
        raise ValueError("compute_optimization_loss: the error maps must hold num_source_imgs * batch pairs")
    maps = [_stacked(fwd_data, inv_data, k) for k in ('diff_img', 'valid_mask', 'auto_mask_error', 'auto_mask', 'weight_mask')]
    w_depth = float(options['l_depth_consist_weight']) if options['l_depth_consist'] == True else 0.0   # noqa: E712
    loss = ops.PftReduceFn.apply(*maps, bsz, n_src, _pft_flags(options), w_depth)
    if options['diff_img_argmin'] != True:   # noqa: E712 -- only the arg-min term keeps the channel dim (optimizer.py:69)
        loss = loss.reshape(())
    if options['l_depth_init'] == True:   # noqa: E712
        loss = loss + options['l_depth_init_weight'] * ops.SsimMeanFn.apply(target_disparity, init_disparity.detach()).reshape(())
    if options['l_smooth'] == True:   # noqa: E712
        loss = loss + options['l_smooth_weight'] * get_smooth_loss(target_disparity, target_img)
    if options['l_pose_consist'] == True:   # noqa: E712
        loss = loss + 0.1 * (fwd_data['poses'] + inv_data['poses']).abs().mean()
    return loss


def compute_photometric_error(target_img, source_img, target_depth, source_depth, pose, intrinsics):
    """helpers.py:8-23: single-pair forward error maps for the loss-surface plots (one warp
    launch + one photometric launch)."""
    img_rec, valid_mask, projected_depth, computed_depth = inverse_warp2(
        source_img, target_depth, source_depth, -pose, intrinsics, 'zeros')
    # auto_mask * valid_mask (helpers.py:18) comes out of the photometric launch itself
    _, diff_img, masked, weight_mask = ops.PhotoErrorFn.apply(
        target_img.detach(), source_img.detach(), img_rec, projected_depth, computed_depth, 0.15, 0.85, valid_mask)
    return {'diff_img': diff_img, 'img_rec': img_rec, 'valid_mask': masked,
            'weight_mask': weight_mask, 'poses': pose}
