"""Drop-in for the reference's ``losses.py``: ``SSIM_Loss``, ``get_smooth_loss`` and
``Compute_Loss`` with the reference's signatures and return structure, computed by
the fused sm_100a kernels.

``Compute_Loss.forward`` submits all 2*S direction/source pair evaluations of a
scale (losses.py:99-127) as ONE forward launch and ONE backward launch; the
mean-on-mask threshold (losses.py:144) is decided on the device, so the loss
never synchronises with the host.
"""
import torch
import torch.nn as nn

from . import _cabi, ops
from .stn import inverse_intrinsics, inverse_intrinsics_forked, inverse_warp2, pose_vec2mat


class SSIM_Loss(nn.Module):
    """Layer to compute the SSIM dissimilarity between a pair of images
    (reference losses.py:11-41): 3x3 mean/variance statistics on reflection-padded
    inputs, clamp((1 - SSIM) / 2, 0, 1), C1 = 0.01**2, C2 = 0.03**2."""

    def __init__(self):
        super(SSIM_Loss, self).__init__()
        self.C1 = 0.01 ** 2
        self.C2 = 0.03 ** 2

    def forward(self, x, y, valid_points=None):
        return ops.SsimFn.apply(x, y)


def disp_to_depth(disp, min_depth, max_depth):
    """utils/learning_helpers.py:77-86."""
    min_disp = 1 / max_depth
    max_disp = 1 / min_depth
    scaled_disp = min_disp + (max_disp - min_disp) * disp
    if torch.is_tensor(disp) and disp.is_cuda and disp.dtype == torch.float32:
        # the reciprocal and its adjoint as one launch each way (PFT calls this for every frame of every epoch,
        # optimization_experiments/optimizer.py:245); same bits as 1 / scaled_disp
        depth = ops.DispToDepthFn.apply(disp, min_disp, max_disp)
    else:
        depth = 1 / scaled_disp
    return scaled_disp, depth


def get_smooth_loss(disp, img):
    """Edge-aware smoothness of the mean-normalised disparity (reference losses.py:43-61), fused
    into four small kernels forward + backward (csrc/smooth_kernels.cu).  Differentiable w.r.t.
    `disp` [B,1,H,W]; the image is data."""
    if img.requires_grad:
        raise NotImplementedError("gradients w.r.t. the image are not implemented "
                                  "(no reference call site differentiates them)")
    if disp.dim() != 4 or disp.size(1) != 1 or img.dim() != 4 or img.size(1) != 3:
        raise ValueError("get_smooth_loss expects disp [B,1,H,W] and img [B,3,H,W]")
    return ops.SmoothLossFn.apply(disp, img)


def _pair_flags(config):
    flags = _cabi.SSIM
    if config['with_auto_mask'] == True:    # noqa: E712 -- the reference compares with ==
        flags |= _cabi.AUTO_MASK
    if config['with_depth_mask']:
        flags |= _cabi.DEPTH_MASK
    if config['l_depth_consist'] == True:   # noqa: E712
        flags |= _cabi.DEPTH_CONSIST
    return flags


class Compute_Loss(nn.modules.Module):
    """Reference losses.py:64-194."""

    def __init__(self, config):
        super(Compute_Loss, self).__init__()
        self.config = config
        self.ssim = SSIM_Loss()
        self.l1_weight = config['l1_weight']
        self.l_ssim_weight = config['l_ssim_weight']
        self.l_smooth_weight = config['l_smooth_weight']
        self.num_scales = config['num_scales']
        self.l_depth_consist_weight = config['l_depth_consist_weight']

    # -- fused evaluation of a list of (tgt_img, ref_img, tgt_depth, ref_depth, pose) pairs
    @staticmethod
    def _refuse_image_grads(specs):
        if any(s[0].requires_grad or s[1].requires_grad for s in specs):
            raise NotImplementedError("gradients w.r.t. the images are not implemented by the fused pair loss "
                                      "(no reference call site differentiates them)")

    def _pair_groups(self, specs, intrinsics):
        """One launch for all `specs`.  Returns per-group (l_reprojection, l_depth,
        diff_img, valid_mask)."""
        if intrinsics.requires_grad:
            raise NotImplementedError("gradients w.r.t. the intrinsics are not implemented")
        self._refuse_image_grads(specs)
        if self.config['l_ssim'] != True:   # noqa: E712
            # L1-only configuration (diff_img keeps its 3 channels, losses.py:154,180): composed from the
            # fused warp; no reference script runs it, so it gets no dedicated kernel
            return [self._pairwise_l1_only(*s, intrinsics) for s in specs]
        n = len(specs)
        kinv = inverse_intrinsics(intrinsics)                        # models/stn.py:257
        poses = torch.cat([s[4][:, 0:6] for s in specs], 0)              # [n*B, 6]
        if poses.is_cuda:
            # models/stn.py:259-262 for every pair in one launch (rounded like the reference's batch-B calls)
            proj = ops.PoseProjFn.apply(poses, intrinsics, 1.0, intrinsics.shape[0])
        else:   # only reachable from the CPU-emulated tests, whose fixtures carry the CPU's sin/cos bits
            proj = torch.cat([intrinsics @ pose_vec2mat(s[4][:, 0:6]) for s in specs], 0)
        tensors = []
        for tgt_img, ref_img, tgt_depth, ref_depth, _ in specs:
            tensors += [tgt_img, ref_img, tgt_depth, ref_depth]
        cfg = (float(self.config['l1_weight']), float(self.config['l_ssim_weight']), _pair_flags(self.config))
        diff, mask, l_rep, l_dep = ops.PairLossFn.apply(cfg, n, kinv, proj, *tensors)
        want_depth = self.config['l_depth_consist'] == True          # noqa: E712
        return [(l_rep[i], l_dep[i] if want_depth else 0, diff[i], mask[i]) for i in range(n)]

    def _pairwise_l1_only(self, tgt_img, ref_img, tgt_depth, ref_depth, pose, intrinsics):
        warped, valid_mask, projected_depth, computed_depth = inverse_warp2(ref_img, tgt_depth, ref_depth, pose, intrinsics)
        diff_img = (tgt_img - warped).abs().clamp(0, 1)
        if self.config['with_auto_mask'] == True:   # noqa: E712
            valid_mask = (diff_img.mean(dim=1, keepdim=True)
                          < (tgt_img - ref_img).abs().mean(dim=1, keepdim=True)).float() * valid_mask
        diff_depth = ((computed_depth - projected_depth).abs() / (computed_depth + projected_depth)).clamp(0, 1)
        if self.config['with_depth_mask']:
            diff_img = diff_img * (1 - diff_depth)
        l_depth = self.mean_on_mask(diff_depth, valid_mask) if self.config['l_depth_consist'] == True else 0   # noqa: E712
        return self.mean_on_mask(diff_img, valid_mask), l_depth, diff_img, valid_mask

    def _can_fuse_frame(self, specs, intrinsics):
        return (self.config['l_ssim'] == True and not intrinsics.requires_grad      # noqa: E712
                and len(specs) <= 8
                and all(s[0].shape == specs[0][0].shape and s[1].shape == specs[0][0].shape for s in specs)
                # disparities [B,1,h,w] of one scale share a resolution (possibly lower than the images')
                and all(s[k].dim() == 4 and s[k].shape == specs[0][2].shape for s in specs for k in (2, 3)))

    def _frame_terms(self, specs, roles, intrinsics, kinv):
        """All pair evaluations of one scale plus disp_to_depth, the pose algebra and the
        min-reprojection / mean-on-mask reductions as one fused autograd node.  `specs` are
        (tgt_img, ref_img, tgt_disp, ref_disp, pose) with the disparities at their own pyramid
        resolution and the un-negated poses.  kinv = None: K^-1 (models/stn.py:257) is computed inside
        the node's first launch.  Returns (terms [3], total [1], kinv): (l_reconstruct_inverse,
        l_reconstruct_forward, l_depth), their sum, and the K^-1 the node used (for the next scale)."""
        images, disps = [], []

        def index(lst, t):
            for i, u in enumerate(lst):
                if u is t:
                    return i
            lst.append(t)
            return len(lst) - 1
        groups = []
        for role, (tgt_img, ref_img, tgt_disp, ref_disp, _) in zip(roles, specs):
            groups.append((0 if role == 'inv' else 1, index(images, tgt_img), index(images, ref_img),
                           index(disps, tgt_disp), index(disps, ref_disp)))
        want_depth = self.config['l_depth_consist'] == True          # noqa: E712
        meta = {"w_l1": float(self.config['l1_weight']), "w_ssim": float(self.config['l_ssim_weight']),
                "flags": _pair_flags(self.config), "w_inverse": 0.3,
                "w_depth": float(self.l_depth_consist_weight) if want_depth else 0.0,
                "min_depth": self.config['min_depth'], "max_depth": self.config['max_depth'],
                "n_img": len(images), "groups": groups}
        # check_sizes accepts [B,8] pose vectors (models/stn.py:252); only the first six enter the warp
        poses = [s[4] if s[4].shape[1] == 6 else s[4][:, 0:6] for s in specs]
        terms, total = ops.FrameLossFn.apply(meta, kinv, intrinsics, *poses, *images, *disps)
        return terms, total, meta["kinv"]

    def forward(self, source_imgs, target_img, poses, disparity, intrinsics, pose_vec_weight=None,
                validate=False, epoch=5, target_img_right=None):
        """Reference losses.py:75-140.  Returns the dict of [1]-shaped tensors
        l_reconstruct_inverse, l_reconstruct_forward, l_depth, l_smooth, total."""
        # (reference: four torch.zeros(1).type_as(intrinsics)); one fill, four one-element slices
        zeros = torch.zeros(4, dtype=intrinsics.dtype, device=intrinsics.device)
        keys = ('l_reconstruct_inverse', 'l_reconstruct_forward', 'l_depth', 'l_smooth')
        losses = {key: zeros[i:i + 1] for i, key in enumerate(keys)}
        fused_total = None              # (inverse + forward) + depth of a single fused scale, from the kernel
        # the kernel's sum is the whole `total` only when this call evaluates exactly one scale and divides
        # by one (the reference loops over every entry of `disparity` whatever num_scales says, losses.py:84,136)
        single_scale = len(disparity[0]) == 1 and self.num_scales == 1
        kinv = None
        disparity, source_disparities = disparity[0], disparity[1:]
        poses, poses_inv = poses[0], poses[1]
        _, _, h, w = target_img.size()
        cfg = self.config
        upsampled = {}

        def full_res(dmap):
            """losses.py:86-87,102-103: the lower scales are nearest-upsampled to the image size.  The fused
            frame node reads them at their own resolution (the upsample is folded into its disp -> depth
            kernel); only the smoothness term and the composed fall-back need the upsampled tensor."""
            if tuple(dmap.shape[-2:]) == (h, w):
                return dmap
            if id(dmap) not in upsampled:
                upsampled[id(dmap)] = nn.functional.interpolate(dmap, (h, w), mode='nearest')
            return upsampled[id(dmap)]

        for scale, disp in enumerate(disparity):
            if cfg['l_smooth']:
                losses['l_smooth'] = losses['l_smooth'] + (self.l_smooth_weight * get_smooth_loss(full_res(disp), target_img)) / (2 ** scale)
            if cfg['l_reconstruction']:
                # (tgt_img, ref_img, tgt_disp, ref_disp, pose); the pose handed to the warp is the
                # negated prediction (losses.py:112,119) -- negated inside the fused node
                specs, roles = [], []
                for j, source_img in enumerate(source_imgs):
                    source_disparity = source_disparities[j][scale]
                    if cfg['l_smooth']:
                        losses['l_smooth'] = losses['l_smooth'] + (self.l_smooth_weight * get_smooth_loss(full_res(source_disparity), source_img)) / (2 ** scale)
                    if cfg['l_inverse']:   # inverse reconstruction: target reprojected into the source frame
                        specs.append((source_img, target_img, source_disparity, disp, poses_inv[j]))
                        roles.append('inv')
                    specs.append((target_img, source_img, disp, source_disparity, poses[j]))
                    roles.append('fwd')
                if self._can_fuse_frame(specs, intrinsics):
                    self._refuse_image_grads(specs)
                    # K^-1 (models/stn.py:257) once per call: the first fused scale computes it inside its prologue launch
                    terms, total, kinv = self._frame_terms(specs, roles, intrinsics, kinv)
                    fresh = scale == 0          # 0 + x == x: skip the add into the zero tensor
                    for i, key in enumerate(keys[:3]):
                        losses[key] = terms[i:i + 1] if fresh else losses[key] + terms[i:i + 1]
                    fused_total = total if single_scale else None
                    continue
                specs = [(s_[0], s_[1], full_res(s_[2]), full_res(s_[3]), s_[4]) for s_ in specs]
                depth_of = {}

                def to_depth(dmap):
                    if id(dmap) not in depth_of:
                        depth_of[id(dmap)] = disp_to_depth(dmap, cfg['min_depth'], cfg['max_depth'])[1]
                    return depth_of[id(dmap)]
                specs = [(s_[0], s_[1], to_depth(s_[2]), to_depth(s_[3]), -s_[4]) for s_ in specs]
                results = self._pair_groups(specs, intrinsics)
                reconstruction_errors = []
                for role, (l_reprojection, l_depth, diff_img, _) in zip(roles, results):
                    if cfg['l_depth_consist']:
                        losses['l_depth'] = losses['l_depth'] + self.l_depth_consist_weight * l_depth
                    if role == 'inv':
                        losses['l_reconstruct_inverse'] = losses['l_reconstruct_inverse'] + 0.3 * l_reprojection
                    else:
                        reconstruction_errors.append(diff_img)
                reconstruction_errors = torch.cat(reconstruction_errors, 1)
                reconstruction_errors, _ = torch.min(reconstruction_errors, 1)
                losses['l_reconstruct_forward'] = losses['l_reconstruct_forward'] + reconstruction_errors.mean()
        if fused_total is not None:
            # total = ((inverse + forward) + depth) + smooth in the reference's order; the first two adds
            # came out of the finalize kernel, x + 0 == x when the smoothness term is off
            losses['total'] = fused_total + losses['l_smooth'] if cfg['l_smooth'] else fused_total
            return losses
        losses['total'] = 0
        for key in keys:
            if self.num_scales != 1:               # x / 1 == x
                losses[key] = losses[key] / (self.num_scales)
            losses['total'] = losses[key] if isinstance(losses['total'], int) else losses['total'] + losses[key]
        return losses

    def mean_on_mask(self, diff, valid_mask):
        """Reference losses.py:142-149, decided on the device (no .item()): the masked
        mean if more than 10000 mask entries are set, else a constant 0."""
        mask = valid_mask.expand_as(diff)
        total = mask.sum()
        mean_value = (diff * mask).sum() / total.clamp(min=1)
        return torch.where(total > 10000, mean_value, torch.zeros_like(mean_value))

    def compute_pairwise_loss(self, tgt_img, ref_img, tgt_depth, ref_depth, pose, intrinsic, epoch, padding_mode='zeros'):
        """Reference losses.py:151-183.  Returns (l_reprojection, l_depth, diff_img,
        valid_mask, None)."""
        if padding_mode != 'zeros':
            raise NotImplementedError("only padding_mode='zeros' is implemented")
        (l_reprojection, l_depth, diff_img, valid_mask), = self._pair_groups(
            [(tgt_img, ref_img, tgt_depth, ref_depth, pose)], intrinsic)
        return l_reprojection, l_depth, diff_img, valid_mask, None

    def compute_reprojection_loss(self, pred, target):
        """Reference losses.py:185-194 (no live caller)."""
        diff_img = torch.abs(target - pred).mean(1, True)
        if self.config['l_ssim'] == True:   # noqa: E712
            ssim_loss = self.ssim(pred, target).mean(1, True)
            diff_img = 0.85 * ssim_loss + 0.15 * diff_img
        return diff_img
