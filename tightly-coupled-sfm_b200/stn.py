"""Drop-in for the reference's ``models/stn.py`` (same names, arguments, return
values and assertion behaviour), with ``inverse_warp2`` running as one fused
sm_100a kernel forward and one backward (csrc/warp_kernels.cu).

Only the small pose -> K[R|t] algebra stays in PyTorch operators, issued in the
reference's order so that K^-1 and the projection matrix carry the reference's
bits and autograd maps grad(K[R|t]) back to the 6-DoF pose.
"""
import torch
import torch.nn.functional as F

from . import ops

pixel_coords = None   # kept for API compatibility (models/stn.py:8); the fused kernel needs no grid


def set_id_grid(depth):
    """models/stn.py:10-21."""
    global pixel_coords
    b, h, w = depth.size()
    i_range = torch.arange(0, h).view(1, h, 1).expand(1, h, w).type_as(depth)
    j_range = torch.arange(0, w).view(1, 1, w).expand(1, h, w).type_as(depth)
    ones = torch.ones(1, h, w).type_as(depth)
    pixel_coords = torch.stack((j_range, i_range, ones), dim=1)


def check_sizes(input, input_name, expected):
    """models/stn.py:24-30 (AssertionError on mismatch; an `expected` given as a
    list of alternatives only checks the number of dimensions, like the reference)."""
    condition = [input.ndimension() == len(expected)]
    for i, size in enumerate(expected):
        if size.isdigit():
            condition.append(input.size(i) == int(size))
    assert (all(condition)), "wrong size for {}, expected {}, got  {}".format(
        input_name, 'x'.join(expected), list(input.size()))


def _axis_rotation(axis, c, s, zeros, ones):
    """[B,3,3] rotation about one coordinate axis from per-sample cos / sin."""
    entries = ([ones, zeros, zeros, zeros, c, -s, zeros, s, c],
               [c, zeros, s, zeros, ones, zeros, -s, zeros, c],
               [c, -s, zeros, s, c, zeros, zeros, zeros, ones])[axis]
    return torch.stack(entries, dim=1).reshape(-1, 3, 3)


def euler2mat(angle):
    """R = Rx @ Ry @ Rz from [B,3] euler angles.  models/stn.py:81-116."""
    zeros = angle[:, 2].detach() * 0
    ones = zeros.detach() + 1
    rx, ry, rz = (_axis_rotation(i, torch.cos(angle[:, i]), torch.sin(angle[:, i]), zeros, ones) for i in range(3))
    return rx @ ry @ rz


def quat2mat(quat):
    """models/stn.py:119-140 (no live caller in the reference)."""
    nq = torch.cat([quat[:, :1].detach() * 0 + 1, quat], dim=1)
    nq = nq / nq.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = nq[:, 0], nq[:, 1], nq[:, 2], nq[:, 3]
    n = quat.size(0)
    w2, x2, y2, z2 = w.pow(2), x.pow(2), y.pow(2), z.pow(2)
    wx, wy, wz = w * x, w * y, w * z
    xy, xz, yz = x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], dim=1).reshape(n, 3, 3)


def pose_vec2mat(vec, rotation_mode='euler'):
    """[R|t] from (tx,ty,tz,rx,ry,rz).  models/stn.py:143-158."""
    translation = vec[:, :3].unsqueeze(-1)
    rot = vec[:, 3:]
    if rotation_mode == 'euler':
        rot_mat = euler2mat(rot)
    elif rotation_mode == 'quat':
        rot_mat = quat2mat(rot)
    return torch.cat([rot_mat, translation], dim=2)


def inverse_intrinsics(intrinsics):
    """`intrinsics.inverse()` (models/stn.py:257), recomputed on every call like the reference does.  On CUDA fp32
    tensors it is ONE launch of the library's own batched 3x3 LU, which reproduces the bits of torch's cuBLAS path
    (getrf + getrs on the identity: ten launches and a layout copy) -- the ordering of that arithmetic was matched bit
    for bit on probed inverses, see csrc/frame_kernels.cu.  It neither synchronises nor breaks CUDA-graph capture.
    Other tensors go through `torch.linalg.inv_ex` (no device -> host error check either)."""
    k = intrinsics.detach()
    if k.is_cuda and k.dtype == torch.float32 and k.dim() == 3:
        from . import _raw
        with ops._guard(k):
            return _raw.intrinsics_inverse(ops.lib(), k)
    return torch.linalg.inv_ex(k)[0]


def inverse_intrinsics_forked(intrinsics):
    """(K^-1, event): kept for callers that overlapped torch's ten-launch inverse with other work on a side stream.  The
    single-launch inverse above costs less than the fork / join did, so it now runs on the caller's stream and the
    event is always None."""
    return inverse_intrinsics(intrinsics).contiguous(), None


def projection_matrices(pose, intrinsics, kinv=None):
    """(K^-1, K @ [R|t]) as models/stn.py:257-262 forms them.  On the GPU the 25-kernel euler/bmm
    chain is one fused launch that reproduces it bit for bit (csrc/frame_kernels.cu), including
    the differently rounded cuBLAS kernel eager PyTorch uses at batch 1."""
    if kinv is None:
        kinv = inverse_intrinsics(intrinsics)
    if pose.is_cuda and not intrinsics.requires_grad:
        return kinv, ops.PoseProjFn.apply(pose[:, 0:6], intrinsics, 1.0)
    return kinv, intrinsics @ pose_vec2mat(pose[:, 0:6])


def inverse_warp2(img, depth, ref_depth, pose, intrinsics, padding_mode='zeros'):
    """Inverse warp a source image to the target image plane (models/stn.py:234-273).

    Args / returns as the reference: img [B,3,H,W] (may be a channel slice of a
    wider stack), depth, ref_depth [B,1,H,W], pose [B,6+], intrinsics [B,3,3] ->
    (projected_img [B,3,H,W], valid_mask [B,1,H,W] float, projected_depth,
    computed_depth [B,1,H,W]).  Differentiable w.r.t. depth, ref_depth, pose and
    (if it requires grad) img.
    """
    check_sizes(img, 'img', 'B3HW')
    check_sizes(depth, 'depth', 'B1HW')
    check_sizes(ref_depth, 'ref_depth', 'B1HW')
    check_sizes(pose, 'pose', ['B6', 'B8'])
    check_sizes(intrinsics, 'intrinsics', 'B33')
    if padding_mode != 'zeros':
        raise NotImplementedError("inverse_warp2: only padding_mode='zeros' is implemented "
                                  "(every reference call site passes 'zeros')")
    if intrinsics.requires_grad:
        raise NotImplementedError("inverse_warp2: gradients w.r.t. the intrinsics are not implemented "
                                  "(no reference call site differentiates them)")
    kinv, proj = projection_matrices(pose, intrinsics)
    return ops.InverseWarp2Fn.apply(img, depth, ref_depth, kinv, proj)


def inverse_warp2_stacked(img, depth, ref_depth, pose, intrinsics, kinv, stack_target, need_depths=True):
    """inverse_warp2 plus the next pose-network input of solve_pose_iteratively
    (train_mono.py:74-76): returns (projected_img, valid_mask, projected_depth, computed_depth,
    stack) with stack = [stack_target * valid_mask | projected_img] as one [B,6,H,W] tensor.
    need_depths=False skips the two depth outputs (returned as None) and their backward."""
    kinv, proj = projection_matrices(pose, intrinsics, kinv)
    return ops.InverseWarp2Fn.apply(img, depth, ref_depth, kinv, proj, stack_target, need_depths)


# ---- legacy helpers kept importable (no live callers in the reference) ----------

def pixel2cam(depth, intrinsics_inv):
    """models/stn.py:33-48."""
    global pixel_coords
    b, h, w = depth.size()
    if (pixel_coords is None) or pixel_coords.size(2) < h or pixel_coords.size(3) < w \
            or pixel_coords.device != depth.device or pixel_coords.dtype != depth.dtype:
        set_id_grid(depth)
    grid = pixel_coords[:, :, :h, :w].expand(b, 3, h, w).reshape(b, 3, -1)
    return (intrinsics_inv @ grid).reshape(b, 3, h, w) * depth.unsqueeze(1)


def _project(cam_coords, proj_c2p_rot, proj_c2p_tr):
    b, _, h, w = cam_coords.size()
    flat = cam_coords.reshape(b, 3, -1)
    pc = proj_c2p_rot @ flat if proj_c2p_rot is not None else flat
    if proj_c2p_tr is not None:
        pc = pc + proj_c2p_tr
    z = pc[:, 2].clamp(min=1e-3)
    return 2 * (pc[:, 0] / z) / (w - 1) - 1, 2 * (pc[:, 1] / z) / (h - 1) - 1, z


def cam2pixel(cam_coords, proj_c2p_rot, proj_c2p_tr, padding_mode):
    """models/stn.py:51-78."""
    b, _, h, w = cam_coords.size()
    xn, yn, _ = _project(cam_coords, proj_c2p_rot, proj_c2p_tr)
    return torch.stack([xn, yn], dim=2).reshape(b, h, w, 2)


def cam2pixel2(cam_coords, proj_c2p_rot, proj_c2p_tr, padding_mode):
    """models/stn.py:198-231."""
    b, _, h, w = cam_coords.size()
    xn, yn, z = _project(cam_coords, proj_c2p_rot, proj_c2p_tr)
    if padding_mode == 'zeros':
        xn = torch.where(((xn > 1) + (xn < -1)).detach(), torch.full_like(xn, 2), xn)
        yn = torch.where(((yn > 1) + (yn < -1)).detach(), torch.full_like(yn, 2), yn)
    return torch.stack([xn, yn], dim=2).reshape(b, h, w, 2), z.reshape(b, 1, h, w)


def inverse_warp(img, depth, pose, intrinsics, rotation_mode='euler', padding_mode='zeros'):
    """models/stn.py:161-195 (legacy two-output variant, no live callers): composed
    from the helpers above."""
    check_sizes(img, 'img', 'B3HW')
    check_sizes(depth, 'depth', 'BHW')
    check_sizes(pose, 'pose', 'B6')
    check_sizes(intrinsics, 'intrinsics', 'B33')
    cam = pixel2cam(depth, intrinsics.inverse())
    proj = intrinsics @ pose_vec2mat(pose, rotation_mode)
    grid = cam2pixel(cam, proj[:, :, :3], proj[:, :, -1:], padding_mode)
    projected = F.grid_sample(img, grid, padding_mode=padding_mode)
    return projected, grid.abs().max(dim=-1)[0] <= 1


def inverse_warp2_op(img, depth, ref_depth, pose, intrinsics, padding_mode='zeros'):
    """inverse_warp2 through the `torch.library` registration (torch.ops.tcsfm.inverse_warp2): the spelling for
    callers that trace with fake tensors / torch.compile; same results as inverse_warp2."""
    from . import custom_ops  # noqa: F401  (registers the operators)
    check_sizes(img, 'img', 'B3HW')
    check_sizes(depth, 'depth', 'B1HW')
    check_sizes(ref_depth, 'ref_depth', 'B1HW')
    check_sizes(pose, 'pose', ['B6', 'B8'])
    check_sizes(intrinsics, 'intrinsics', 'B33')
    if padding_mode != 'zeros':
        raise NotImplementedError("only padding_mode='zeros' is implemented")
    kinv = inverse_intrinsics(intrinsics).contiguous()
    proj = intrinsics @ pose_vec2mat(pose[:, 0:6])
    return torch.ops.tcsfm.inverse_warp2(img, depth.contiguous(), ref_depth.contiguous(), kinv, proj.contiguous())
