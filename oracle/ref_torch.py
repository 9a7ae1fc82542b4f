"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not shipped, not imported by the product.

Op-for-op restatement, in stock PyTorch (ATen) operators, of the reference's
differentiable inverse-warp + SSIM/L1 photometric-loss path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this file, and only as the checker or as
the timed CPU baseline -- never as (part of) the product path.

Why it is written with torch operators: the reference *is* eager PyTorch, its
arithmetic lives in ATen (torch 2.11.0+cu128 here; the reference pins no
version, README.md:8-14).  Issuing the same ATen operators in the same order
makes this file bit-identical to the reference on the same device, which is
what "validity masks bit-exact" needs.  It is device
agnostic: on ``cpu`` it is the CPU baseline, on ``cuda`` it is the reference's
eager GPU path (the same-device arbiter for bit-level comparisons).

Parity pinning: the reference ships no tests or golden vectors (SURVEY.md §4,
§8c).  This restatement is pinned against outputs of the reference itself,
imported in the build container: ``tests/golden/make_golden.py`` generates the
committed fixtures from ``/root/reference``; ``tests/test_oracle.py`` checks
this file bit-for-bit against those fixtures (and against the live reference
when it is present).

Every function cites the reference lines it follows (paths are relative to the
reference root).
"""
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# geometry (models/stn.py)
# ----------------------------------------------------------------------------

def euler_to_matrix(angle):
    """R = Rx(a0) @ Ry(a1) @ Rz(a2).  models/stn.py:81-116
    (duplicate: utils/geometry_helpers.py:5-40)."""
    n = angle.size(0)
    ax, ay, az = angle[:, 0], angle[:, 1], angle[:, 2]
    zero = az.detach() * 0
    one = zero.detach() + 1
    cz, sz = torch.cos(az), torch.sin(az)
    rz = torch.stack([cz, -sz, zero, sz, cz, zero, zero, zero, one], dim=1).reshape(n, 3, 3)
    cy, sy = torch.cos(ay), torch.sin(ay)
    ry = torch.stack([cy, zero, sy, zero, one, zero, -sy, zero, cy], dim=1).reshape(n, 3, 3)
    cx, sx = torch.cos(ax), torch.sin(ax)
    rx = torch.stack([one, zero, zero, zero, cx, -sx, zero, sx, cx], dim=1).reshape(n, 3, 3)
    return rx @ ry @ rz


def pose_to_matrix(vec):
    """[R | t] from (tx,ty,tz,rx,ry,rz).  models/stn.py:143-158 (euler branch)."""
    return torch.cat([euler_to_matrix(vec[:, 3:]), vec[:, :3].unsqueeze(-1)], dim=2)


def pixel_grid(h, w, like):
    """Homogeneous pixel grid [1,3,H,W] = (col, row, 1).  models/stn.py:10-21."""
    rows = torch.arange(0, h).view(1, h, 1).expand(1, h, w).type_as(like)
    cols = torch.arange(0, w).view(1, 1, w).expand(1, h, w).type_as(like)
    return torch.stack((cols, rows, torch.ones(1, h, w).type_as(like)), dim=1)


def backproject(depth, k_inv):
    """cam = depth * (K^-1 @ [u,v,1]).  depth [B,H,W].  models/stn.py:33-48
    (the reference's module-global grid cache is replaced by a fresh grid; the
    values are identical)."""
    b, h, w = depth.size()
    grid = pixel_grid(h, w, depth).expand(b, 3, h, w).reshape(b, 3, -1)
    return (k_inv @ grid).reshape(b, 3, h, w) * depth.unsqueeze(1)


def project_normalised(cam, rot, tr, padding_mode="zeros"):
    """Normalised sampling grid + clamped depth.  models/stn.py:198-231."""
    b, _, h, w = cam.size()
    pc = rot @ cam.reshape(b, 3, -1)
    pc = pc + tr
    x, y = pc[:, 0], pc[:, 1]
    z = pc[:, 2].clamp(min=1e-3)
    xn = 2 * (x / z) / (w - 1) - 1
    yn = 2 * (y / z) / (h - 1) - 1
    if padding_mode == "zeros":
        # reference: X_norm[X_mask] = 2 (in place, mask detached) -- same
        # values and the same (zero) gradient on the overwritten entries
        xm = ((xn > 1) + (xn < -1)).detach()
        xn = torch.where(xm, torch.full_like(xn, 2), xn)
        ym = ((yn > 1) + (yn < -1)).detach()
        yn = torch.where(ym, torch.full_like(yn, 2), yn)
    grid = torch.stack([xn, yn], dim=2)
    return grid.reshape(b, h, w, 2), z.reshape(b, 1, h, w)


def inverse_warp2(img, depth, ref_depth, pose, intrinsics, padding_mode="zeros"):
    """models/stn.py:234-273.  Returns (projected_img, valid_mask,
    projected_depth, computed_depth)."""
    assert img.dim() == 4 and img.size(1) == 3
    assert depth.dim() == 4 and depth.size(1) == 1
    assert ref_depth.dim() == 4 and ref_depth.size(1) == 1
    assert pose.dim() == 2 and intrinsics.dim() == 3
    cam = backproject(depth.squeeze(1), intrinsics.inverse())
    proj = intrinsics @ pose_to_matrix(pose[:, 0:6])
    rot, tr = proj[:, :, :3], proj[:, :, -1:]
    grid, computed_depth = project_normalised(cam, rot, tr, padding_mode)
    projected_img = F.grid_sample(img, grid, padding_mode=padding_mode, align_corners=False)
    valid = (grid.abs().max(dim=-1)[0] <= 1).unsqueeze(1).float()
    projected_depth = F.grid_sample(ref_depth, grid, padding_mode=padding_mode, align_corners=False)
    return projected_img, valid, projected_depth, computed_depth


# ----------------------------------------------------------------------------
# photometric terms (losses.py)
# ----------------------------------------------------------------------------

SSIM_C1 = 0.01 ** 2
SSIM_C2 = 0.03 ** 2


def ssim_dissimilarity(x, y):
    """3x3 reflect-padded SSIM dissimilarity map.  losses.py:11-41."""
    x = F.pad(x, (1, 1, 1, 1), mode="reflect")
    y = F.pad(y, (1, 1, 1, 1), mode="reflect")
    mu_x = F.avg_pool2d(x, 3, 1)
    mu_y = F.avg_pool2d(y, 3, 1)
    sigma_x = F.avg_pool2d(x ** 2, 3, 1) - mu_x ** 2
    sigma_y = F.avg_pool2d(y ** 2, 3, 1) - mu_y ** 2
    sigma_xy = F.avg_pool2d(x * y, 3, 1) - mu_x * mu_y
    num = (2 * mu_x * mu_y + SSIM_C1) * (2 * sigma_xy + SSIM_C2)
    den = (mu_x ** 2 + mu_y ** 2 + SSIM_C1) * (sigma_x + sigma_y + SSIM_C2)
    return torch.clamp((1 - num / den) / 2, 0, 1)


def smooth_loss(disp, img):
    """Edge-aware smoothness of the mean-normalised disparity.  losses.py:43-61."""
    mean_disp = disp.mean(2, True).mean(3, True)
    disp = disp / (mean_disp + 1e-7)
    gdx = torch.abs(disp[:, :, :, :-1] - disp[:, :, :, 1:])
    gdy = torch.abs(disp[:, :, :-1, :] - disp[:, :, 1:, :])
    gix = torch.mean(torch.abs(img[:, :, :, :-1] - img[:, :, :, 1:]), 1, keepdim=True)
    giy = torch.mean(torch.abs(img[:, :, :-1, :] - img[:, :, 1:, :]), 1, keepdim=True)
    gdx = gdx * torch.exp(-gix)
    gdy = gdy * torch.exp(-giy)
    return gdx.mean() + gdy.mean()


def disp_to_depth(disp, min_depth, max_depth):
    """utils/learning_helpers.py:77-86."""
    min_disp = 1 / max_depth
    max_disp = 1 / min_depth
    scaled = min_disp + (max_disp - min_disp) * disp
    return scaled, 1 / scaled


def mean_on_mask(diff, valid_mask):
    """losses.py:142-149 (including the host-side branch on the mask sum)."""
    mask = valid_mask.expand_as(diff)
    if mask.sum() > 10000:
        return (diff * mask).sum() / mask.sum()
    return torch.tensor(0).float().type_as(mask)


DEFAULT_LOSS_CONFIG = {
    # run_mono_training.py:27,50-64 defaults
    "l1_weight": 0.15, "l_ssim_weight": 0.85, "l_smooth_weight": 0.05,
    "num_scales": 1, "l_depth_consist_weight": 0.14,
    "min_depth": 0.06, "max_depth": 2.67,
    "l_smooth": True, "l_reconstruction": True, "l_inverse": True,
    "l_depth_consist": False, "with_auto_mask": True, "l_ssim": True,
    "with_depth_mask": False,
}


def pairwise_loss(cfg, tgt_img, ref_img, tgt_depth, ref_depth, pose, intrinsic, padding_mode="zeros"):
    """losses.py:151-183.  Returns (l_reprojection, l_depth, diff_img,
    valid_mask, None)."""
    warped, valid_mask, projected_depth, computed_depth = inverse_warp2(
        ref_img, tgt_depth, ref_depth, pose, intrinsic, padding_mode)
    diff_img = (tgt_img - warped).abs().clamp(0, 1)
    if cfg["with_auto_mask"] == True:  # noqa: E712  (reference compares with ==)
        auto = (diff_img.mean(dim=1, keepdim=True)
                < (tgt_img - ref_img).abs().mean(dim=1, keepdim=True)).float() * valid_mask
        valid_mask = auto
    if cfg["l_ssim"] == True:  # noqa: E712
        ssim_map = ssim_dissimilarity(tgt_img, warped)
        diff_img = (cfg["l1_weight"] * diff_img + cfg["l_ssim_weight"] * ssim_map).mean(1, True)
    l_depth = 0
    diff_depth = ((computed_depth - projected_depth).abs()
                  / (computed_depth + projected_depth)).clamp(0, 1)
    if cfg["with_depth_mask"]:
        diff_img = diff_img * (1 - diff_depth.clone())
    if cfg["l_depth_consist"] == True:  # noqa: E712
        l_depth = mean_on_mask(diff_depth, valid_mask)
    l_reprojection = mean_on_mask(diff_img, valid_mask)
    return l_reprojection, l_depth, diff_img, valid_mask, None


def compute_loss(cfg, source_imgs, target_img, poses, disparity, intrinsics):
    """Compute_Loss.forward, losses.py:75-140.  Returns the dict of [1]-shaped
    tensors ``l_reconstruct_inverse, l_reconstruct_forward, l_depth, l_smooth,
    total``."""
    zero = torch.zeros(1).type_as(intrinsics)
    out = {k: zero.clone() for k in
           ("l_reconstruct_inverse", "l_reconstruct_forward", "l_depth", "l_smooth")}
    tgt_disps, src_disps = disparity[0], disparity[1:]
    fwd_poses, inv_poses = poses[0], poses[1]
    _, _, h, w = target_img.size()
    for scale, disp in enumerate(tgt_disps):
        if scale != 0:
            disp = F.interpolate(disp, (h, w), mode="nearest")
        _, d = disp_to_depth(disp, cfg["min_depth"], cfg["max_depth"])
        if cfg["l_smooth"]:
            out["l_smooth"] += (cfg["l_smooth_weight"] * smooth_loss(disp, target_img)) / (2 ** scale)
        errors = []
        if cfg["l_reconstruction"]:
            for j, src in enumerate(source_imgs):
                pose, pose_inv = fwd_poses[j], inv_poses[j]
                sdisp = src_disps[j][scale]
                if scale != 0:
                    sdisp = F.interpolate(sdisp, (h, w), mode="nearest")
                _, sd = disp_to_depth(sdisp, cfg["min_depth"], cfg["max_depth"])
                if cfg["l_smooth"]:
                    out["l_smooth"] += (cfg["l_smooth_weight"] * smooth_loss(sdisp, src)) / (2 ** scale)
                if cfg["l_inverse"]:
                    l_rep, l_dep, _, _, _ = pairwise_loss(cfg, src, target_img, sd, d, -pose_inv.clone(), intrinsics)
                    if cfg["l_depth_consist"]:
                        out["l_depth"] += cfg["l_depth_consist_weight"] * l_dep
                    out["l_reconstruct_inverse"] += 0.3 * l_rep
                l_rep, l_dep, diff_img, _, _ = pairwise_loss(cfg, target_img, src, d, sd, -pose.clone(), intrinsics)
                if cfg["l_depth_consist"]:
                    out["l_depth"] += cfg["l_depth_consist_weight"] * l_dep
                errors.append(diff_img)
            errors = torch.cat(errors, 1)
            errors, _ = torch.min(errors, 1)
            out["l_reconstruct_forward"] += errors.mean()
    total = 0
    for key in list(out.keys()):
        out[key] = out[key] / cfg["num_scales"]
        total = total + out[key]
    out["total"] = total
    return out


# ----------------------------------------------------------------------------
# iterative egomotion coupling and the PFT error maps (train_mono.py)
# ----------------------------------------------------------------------------

def pft_error_maps(imgs, img_rec, projected_depth, computed_depth):
    """train_mono.py:84-92.  ``imgs`` is the 6-channel [recon-target | source]
    stack.  Returns (auto_mask_error, diff_img, auto_mask, weight_mask)."""
    tgt, src = imgs[:, 0:3], imgs[:, 3:6]
    auto_err = (0.15 * (tgt - src).abs().clamp(0, 1) + 0.85 * ssim_dissimilarity(tgt, src)).mean(1, True)
    diff = (0.15 * (img_rec - tgt.clone().detach()).abs().clamp(0, 1)
            + 0.85 * ssim_dissimilarity(tgt.clone().detach(), img_rec)).mean(1, True)
    auto_mask = (diff < auto_err).float()
    diff_depth = ((computed_depth - projected_depth).abs() / (computed_depth + projected_depth)).clamp(0, 1)
    return auto_err, diff, auto_mask, 1 - diff_depth


def iterative_pose(num_iter, depths, pose_model, target_img, source_img_list, intrinsics, return_errors=False):
    """solve_pose_iteratively, train_mono.py:41-120."""
    n_src = len(source_img_list)
    bsz = target_img.shape[0]
    split = n_src * bsz
    depth, source_depths = depths[0], torch.cat(depths[1:], 0)
    target_depths = depth.repeat(n_src, 1, 1, 1)
    source_imgs = torch.cat(source_img_list, 0)
    intrinsics = intrinsics.repeat(2 * n_src, 1, 1)
    target_imgs = target_img.repeat(n_src, 1, 1, 1)
    imgs = torch.cat([torch.cat([target_imgs, source_imgs], 1),
                      torch.cat([source_imgs, target_imgs], 1)], 0)
    full_poses = pose_model(imgs)
    tgt_depth_full = torch.cat([target_depths, source_depths], 0)
    src_depth_full = torch.cat([source_depths, target_depths], 0)
    img_rec, valid_mask, proj_d, comp_d = inverse_warp2(
        imgs[:, 3:6], tgt_depth_full, src_depth_full, -full_poses, intrinsics, "zeros")
    stacked = full_poses.clone().unsqueeze(1)
    for _ in range(0, num_iter - 1):
        new_imgs = imgs.clone()
        new_imgs[:, 0:3] = new_imgs[:, 0:3] * valid_mask
        new_imgs[:, 3:6] = img_rec
        full_poses = full_poses + pose_model(new_imgs)
        stacked = torch.cat([stacked, full_poses.clone().unsqueeze(1)], 1)
        img_rec, valid_mask, proj_d, comp_d = inverse_warp2(
            imgs[:, 3:6], tgt_depth_full, src_depth_full, -full_poses, intrinsics, "zeros")
    outputs = {"fwd": {}, "inv": {}}
    if return_errors:
        auto_err, diff, auto_mask, weight = pft_error_maps(imgs, img_rec, proj_d, comp_d)
        for name, sl in (("fwd", slice(0, split)), ("inv", slice(split, None))):
            outputs[name] = {"diff_img": diff[sl], "img_rec": img_rec[sl], "valid_mask": valid_mask[sl],
                             "weight_mask": weight[sl], "poses": stacked[sl],
                             "auto_mask_error": auto_err[sl], "auto_mask": auto_mask[sl]}
        new_imgs = imgs.clone()
        new_imgs[:, 0:3] = new_imgs[:, 0:3] * valid_mask
        new_imgs[:, 3:6] = img_rec
        outputs["comb"] = {"imgs": new_imgs, "valid_mask": valid_mask}
    fwd, inv = stacked[0:split].clone(), stacked[split:].clone()
    poses = [fwd[bsz * i:bsz * (i + 1)][:, -1] for i in range(n_src)]
    poses_inv = [inv[bsz * i:bsz * (i + 1)][:, -1] for i in range(n_src)]
    if return_errors:
        return poses, poses_inv, outputs
    return poses, poses_inv


# ----------------------------------------------------------------------------
# PFT window loss (optimization_experiments/optimizer.py, helpers.py)
# ----------------------------------------------------------------------------

DEFAULT_PFT_OPTIONS = {
    # optimization_experiments/run_sequential_optimization.py:69-99 (loss-relevant keys)
    "num_source_imgs": 2, "diff_img_argmin": True, "automasking": True,
    "l_inverse_reconstruction": True, "l_depth_consist": True, "l_depth_consist_weight": 0.15,
    "l_depth_init": True, "l_depth_init_weight": 0.1, "l_smooth": False, "l_smooth_weight": 0.05,
    "l_pose_consist": False, "plotting": False,
}


def pft_window_loss(options, target_img, target_disparity, init_disparity, fwd, inv):
    """DepthOptimizer.compute_optimization_loss, optimizer.py:29-97 (plotting
    branches omitted: they do not touch the loss)."""
    bsz = target_img.shape[0]
    n_src = options["num_source_imgs"]
    loss = 0
    if options["diff_img_argmin"] == True:  # noqa: E712
        stack = torch.cat([fwd["diff_img"][i * bsz:(i + 1) * bsz] for i in range(n_src)], 1).unsqueeze(2)
        diff_min, _ = torch.min(stack, 1)
        vmask = torch.cat([fwd["valid_mask"][i * bsz:(i + 1) * bsz] for i in range(n_src)], 1)
        vmask = vmask.sum(1, keepdim=True).clamp(0, 1)
        if options["automasking"] == True:  # noqa: E712
            aerr = torch.cat([fwd["auto_mask_error"][i * bsz:(i + 1) * bsz] for i in range(n_src)], 1).unsqueeze(2)
            amin, _ = torch.min(aerr, 1)
            vmask = (diff_min < amin).float() * vmask
        loss += (diff_min * vmask * fwd["weight_mask"][0:bsz]).sum(3).sum(2).sum(0) / vmask.sum(3).sum(2).sum(0)
    masked = fwd["diff_img"] * fwd["valid_mask"] * fwd["weight_mask"]
    if options["diff_img_argmin"] == False:  # noqa: E712
        loss += 0.25 * masked.sum() / fwd["valid_mask"].sum()
    masked_inv = inv["diff_img"] * inv["valid_mask"] * inv["weight_mask"]
    if options["l_inverse_reconstruction"] == True:  # noqa: E712
        if options["automasking"] == True:  # noqa: E712
            masked_inv = masked_inv * inv["auto_mask"]
            loss += 0.25 * masked_inv.sum() / (inv["valid_mask"] * inv["auto_mask"]).sum()
        else:
            loss += 0.25 * masked_inv.sum() / inv["valid_mask"].sum()
    if options["l_depth_consist"] == True:  # noqa: E712
        loss += options["l_depth_consist_weight"] * ((-fwd["weight_mask"] + 1)).mean()
        if options["l_inverse_reconstruction"] == True:  # noqa: E712
            loss += options["l_depth_consist_weight"] * ((-inv["weight_mask"] + 1)).mean()
    if options["l_depth_init"] == True:  # noqa: E712
        loss += options["l_depth_init_weight"] * ssim_dissimilarity(
            target_disparity, init_disparity.clone().detach()).mean()
    if options["l_smooth"] == True:  # noqa: E712
        loss += options["l_smooth_weight"] * smooth_loss(target_disparity, target_img)
    if options["l_pose_consist"] == True:  # noqa: E712
        loss += 0.1 * (fwd["poses"] + inv["poses"]).abs().mean()
    return loss


def photometric_error(target_img, source_img, target_depth, source_depth, pose, intrinsics):
    """compute_photometric_error, optimization_experiments/helpers.py:8-23."""
    img_rec, valid_mask, proj_d, comp_d = inverse_warp2(
        source_img, target_depth, source_depth, -pose, intrinsics, "zeros")
    tgt = target_img.clone().detach()
    diff = (0.15 * (img_rec - tgt).abs().clamp(0, 1) + 0.85 * ssim_dissimilarity(tgt, img_rec)).mean(1, True)
    diff_depth = ((comp_d - proj_d).abs() / (comp_d + proj_d)).clamp(0, 1)
    auto = (0.15 * (source_img - tgt).abs().clamp(0, 1) + 0.85 * ssim_dissimilarity(tgt, source_img)).mean(1, True)
    auto = (diff < auto).float()
    return {"diff_img": diff, "img_rec": img_rec, "valid_mask": auto * valid_mask,
            "weight_mask": 1 - diff_depth, "poses": pose}
