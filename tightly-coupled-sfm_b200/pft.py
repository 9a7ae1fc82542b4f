"""Drop-ins for the loss half of the reference's per-frame test-time optimisation
(PFT): ``DepthOptimizer.compute_optimization_loss`` (optimization_experiments/
optimizer.py:29-97) and ``compute_photometric_error`` (optimization_experiments/
helpers.py:8-23).  The optimisation driver itself (deep copies of the depth net,
Adam, plotting; optimizer.py:136-297) is the reference's harness and calls these.
"""
import torch

from . import ops
from .losses import SSIM_Loss, get_smooth_loss
from .stn import inverse_warp2


def compute_optimization_loss(options, target_img, target_disparity, init_disparity, fwd_data, inv_data,
                              ssim_loss=None):
    """optimizer.py:45-97 without the plotting branches.  `init_disparity` is the
    reference's ``self.target_disparity`` (the un-optimised prediction)."""
    ssim_loss = ssim_loss or SSIM_Loss()
    bsz = target_img.shape[0]
    n_src = options['num_source_imgs']
    loss = 0
    if options['diff_img_argmin'] == True:   # noqa: E712
        stack = torch.cat([fwd_data['diff_img'][i * bsz:(i + 1) * bsz] for i in range(n_src)], 1).unsqueeze(2)
        diff_min, _ = torch.min(stack, 1)
        vmask = torch.cat([fwd_data['valid_mask'][i * bsz:(i + 1) * bsz] for i in range(n_src)], 1)
        vmask = vmask.sum(1, keepdim=True).clamp(0, 1)
        if options['automasking'] == True:   # noqa: E712
            aerr = torch.cat([fwd_data['auto_mask_error'][i * bsz:(i + 1) * bsz] for i in range(n_src)], 1).unsqueeze(2)
            amin, _ = torch.min(aerr, 1)
            vmask = (diff_min < amin).float() * vmask
        loss += (diff_min * vmask * fwd_data['weight_mask'][0:bsz]).sum(3).sum(2).sum(0) / vmask.sum(3).sum(2).sum(0)
    masked = fwd_data['diff_img'] * fwd_data['valid_mask'] * fwd_data['weight_mask']
    if options['diff_img_argmin'] == False:   # noqa: E712
        loss += 0.25 * masked.sum() / fwd_data['valid_mask'].sum()
    masked_inv = inv_data['diff_img'] * inv_data['valid_mask'] * inv_data['weight_mask']
    if options['l_inverse_reconstruction'] == True:   # noqa: E712
        if options['automasking'] == True:   # noqa: E712
            masked_inv = masked_inv * inv_data['auto_mask']
            loss += 0.25 * masked_inv.sum() / (inv_data['valid_mask'] * inv_data['auto_mask']).sum()
        else:
            loss += 0.25 * masked_inv.sum() / inv_data['valid_mask'].sum()
    if options['l_depth_consist'] == True:   # noqa: E712
        loss += options['l_depth_consist_weight'] * ((-fwd_data['weight_mask'] + 1)).mean()
        if options['l_inverse_reconstruction'] == True:   # noqa: E712
            loss += options['l_depth_consist_weight'] * ((-inv_data['weight_mask'] + 1)).mean()
    if options['l_depth_init'] == True:   # noqa: E712
        loss += options['l_depth_init_weight'] * ssim_loss(target_disparity, init_disparity.clone().detach()).mean()
    if options['l_smooth'] == True:   # noqa: E712
        loss += options['l_smooth_weight'] * get_smooth_loss(target_disparity, target_img)
    if options['l_pose_consist'] == True:   # noqa: E712
        loss += 0.1 * (fwd_data['poses'] + inv_data['poses']).abs().mean()
    return loss


def compute_photometric_error(target_img, source_img, target_depth, source_depth, pose, intrinsics):
    """helpers.py:8-23: single-pair forward error maps for the loss-surface plots (one warp
    launch + one photometric launch)."""
    img_rec, valid_mask, projected_depth, computed_depth = inverse_warp2(
        source_img, target_depth, source_depth, -pose, intrinsics, 'zeros')
    _, diff_img, auto_mask, weight_mask = ops.PhotoErrorFn.apply(
        target_img.detach(), source_img.detach(), img_rec, projected_depth, computed_depth, 0.15, 0.85)
    return {'diff_img': diff_img, 'img_rec': img_rec, 'valid_mask': auto_mask * valid_mask,
            'weight_mask': weight_mask, 'poses': pose}
