"""Drop-in for the hot-path half of the reference's ``train_mono.py``: the iterative
egomotion coupling ``solve_pose_iteratively`` (pose-net <-> warp <-> pose-net ...)
and its ``return_errors`` photometric block (train_mono.py:41-120), plus the small
helpers around it.  The pose network is whatever callable the caller hands in (the
reference's PyTorch models are out of scope and run unmodified).
"""
import torch

from . import ops
from .losses import SSIM_Loss
from .stn import inverse_intrinsics, inverse_warp2_stacked


def compute_pose_consistency_loss(poses, poses_inv):
    """train_mono.py:8-16: mean |pose + pose_inv| summed over the sources."""
    total = 0
    for pose, pose_inv in zip(poses, poses_inv):
        total += (pose[:, 0:6] + pose_inv[:, 0:6]).abs()
    return total.mean()


def solve_pose(pose_model, target_img, source_img_list, flow_imgs):
    """train_mono.py:18-39 (single-shot pose prediction in both directions)."""
    poses, poses_inv = [], []
    flow_fwd, flow_back = flow_imgs
    for source_img, f_fwd, f_back in zip(source_img_list, flow_fwd, flow_back):
        fwd, back = [target_img, source_img], [source_img, target_img]
        if flow_fwd[0] != None:   # noqa: E711 -- as in the reference
            fwd.append(f_fwd)
            back.append(f_back)
        poses.append(pose_model(torch.cat(fwd, 1)))
        poses_inv.append(pose_model(torch.cat(back, 1)))
    return poses, poses_inv


def solve_disp(depth_model, target_img, source_img_list):
    """train_mono.py:122-132: one depth-net pass over [target, source_1, source_2]."""
    n = target_img.shape[0]
    disparities = depth_model(torch.cat([target_img] + source_img_list, 0))
    return [[d[0:n] for d in disparities], [d[n:2 * n] for d in disparities], [d[2 * n:3 * n] for d in disparities]]


def photometric_error_maps(imgs, img_rec, projected_depth, computed_depth, ssim_loss=None):
    """The ``return_errors`` arithmetic of train_mono.py:84-92 for a stack of pairs, as one
    fused launch (csrc/photo_kernels.cu).  imgs is the 6-channel [reconstruction target |
    source] stack (data: it carries no gradient in any reference call site).  Returns
    (auto_mask_error, diff_img, auto_mask, weight_mask)."""
    if imgs.requires_grad:
        raise NotImplementedError("gradients w.r.t. the input images are not implemented "
                                  "(no reference call site differentiates them)")
    return ops.PhotoErrorFn.apply(imgs[:, 0:3], imgs[:, 3:6], img_rec, projected_depth, computed_depth, 0.15, 0.85)


def solve_pose_iteratively(num_iter, depths, pose_model, target_img, source_img_list, intrinsics, return_errors=False):
    """train_mono.py:41-120.  Stacks the forward (target<-source) and inverse
    (source<-target) pairs of all S sources into one batch of 2*S*B, then alternates
    pose_model and the fused inverse warp `num_iter` times.  Returns (poses,
    poses_inv[, outputs]) exactly like the reference."""
    n_src = len(source_img_list)
    bsz = target_img.shape[0]
    split = n_src * bsz
    depth, source_depths = depths[0], torch.cat(depths[1:], 0)
    target_depths = depth.repeat(n_src, 1, 1, 1)
    source_imgs = torch.cat(source_img_list, 0)
    intrinsics_in = intrinsics
    intrinsics = intrinsics.repeat(2 * n_src, 1, 1)
    target_imgs = target_img.repeat(n_src, 1, 1, 1)
    imgs = torch.cat([torch.cat([target_imgs, source_imgs], 1), torch.cat([source_imgs, target_imgs], 1)], 0)
    tgt_depth_full = torch.cat([target_depths, source_depths], 0)
    src_depth_full = torch.cat([source_depths, target_depths], 0)

    kinv = inverse_intrinsics(intrinsics_in).repeat(2 * n_src, 1, 1)
    tgt_view, src_view = imgs[:, 0:3], imgs[:, 3:6]

    def warp(poses, last):
        # one launch: the reconstruction, its masks/depths and the next pose-net input
        # [target * valid | reconstruction] (train_mono.py:69-76,80).  projected_depth / computed_depth are
        # consumed by the return_errors block of the LAST iteration only (train_mono.py:91): the other
        # iterations neither compute nor store them (8 B/px each way)
        return inverse_warp2_stacked(src_view, tgt_depth_full, src_depth_full, -poses, intrinsics, kinv, tgt_view,
                                     need_depths=return_errors and last)

    full_poses = pose_model(imgs)
    stacked = [full_poses.clone()]
    img_rec, valid_mask, proj_d, comp_d, new_imgs = warp(full_poses, num_iter == 1)
    for it in range(num_iter - 1):
        full_poses = full_poses + pose_model(new_imgs)
        stacked.append(full_poses.clone())
        img_rec, valid_mask, proj_d, comp_d, new_imgs = warp(full_poses, it == num_iter - 2)
    stacked = torch.stack(stacked, 1)                       # [2*S*B, num_iter, 6]

    outputs = {'fwd': {}, 'inv': {}}
    if return_errors:
        auto_err, diff, auto_mask, weight = photometric_error_maps(imgs, img_rec, proj_d, comp_d)
        for name, sl in (('fwd', slice(0, split)), ('inv', slice(split, None))):
            outputs[name] = {'diff_img': diff[sl], 'img_rec': img_rec[sl], 'valid_mask': valid_mask[sl],
                             'weight_mask': weight[sl], 'poses': stacked[sl],
                             'auto_mask_error': auto_err[sl], 'auto_mask': auto_mask[sl]}
        outputs['comb'] = {'imgs': new_imgs, 'valid_mask': valid_mask}

    last = stacked[:, -1]
    poses = [last[bsz * i:bsz * (i + 1)] for i in range(n_src)]
    poses_inv = [last[split + bsz * i:split + bsz * (i + 1)] for i in range(n_src)]
    if return_errors:
        return poses, poses_inv, outputs
    return poses, poses_inv
