"""Synthetic-data driver for per-frame test-time optimisation (PFT), mirroring the loop of
the reference's `DepthOptimizer.optimize_window` (optimization_experiments/optimizer.py:
136-297) around the fused hot path: per window minibatch a private copy of the depth
network is optimised for `epochs` steps of  depth-net -> disp_to_depth ->
solve_pose_iteratively(return_errors=True) -> compute_optimization_loss -> backward -> Adam.

The reference's networks, loaders, ScaleRecovery, plotting and trajectory stitching are out
of scope; the stand-in networks (synth.TinyDepthNet / TinyPoseNet) only give the loop
something differentiable to drive.  Windows are independent, so a sequence is sharded across
ranks with `shard.shard_range` and no communication.
"""
import copy

import torch

from . import losses, pft, train_mono

DEFAULT_OPTIONS = {
    # optimization_experiments/run_sequential_optimization.py:69-99 (loss-relevant keys)
    "num_source_imgs": 2, "diff_img_argmin": True, "automasking": True, "l_inverse_reconstruction": True,
    "l_depth_consist": True, "l_depth_consist_weight": 0.15, "l_depth_init": True, "l_depth_init_weight": 0.1,
    "l_smooth": False, "l_smooth_weight": 0.05, "l_pose_consist": False, "plotting": False,
    "epochs": 20, "lr": 2e-4,
}


class Backend:
    """The three hot-path callables the loop needs; tests swap in the oracle's."""
    solve_pose_iteratively = staticmethod(train_mono.solve_pose_iteratively)
    compute_optimization_loss = staticmethod(pft.compute_optimization_loss)
    disp_to_depth = staticmethod(losses.disp_to_depth)


def optimize_window(depth_net, pose_net, target_img, source_imgs, intrinsics, options=None, iterations=4,
                    depth_range=(0.06, 2.67), backend=Backend):
    """One window minibatch (optimizer.py:136-297).  Returns dict(losses=[per-epoch loss tensors],
    disparity=final target disparity, poses=..., poses_inv=...)."""
    opts = dict(DEFAULT_OPTIONS, **(options or {}))
    bsz = target_img.shape[0]
    imgs = torch.cat([target_img] + list(source_imgs), 0)
    with torch.no_grad():                                    # un-optimised prediction, optimizer.py:143-160
        init_disp = depth_net(imgs)[0][0:bsz].clone()
    net = copy.deepcopy(depth_net)                            # optimizer.py:177-182
    optim = torch.optim.Adam(net.encoder.parameters(), lr=opts["lr"])
    loss_log, out = [], {}
    for epoch in range(opts["epochs"]):                       # optimizer.py:217-268
        optim.zero_grad(set_to_none=True)
        disp = net(imgs)[0]
        disps = [disp[i * bsz:(i + 1) * bsz] for i in range(1 + len(source_imgs))]
        depths = [backend.disp_to_depth(d, depth_range[0], depth_range[1])[1] for d in disps]
        poses, poses_inv, outputs = backend.solve_pose_iteratively(
            iterations, depths, pose_net, target_img, list(source_imgs), intrinsics, return_errors=True)
        loss = backend.compute_optimization_loss(opts, target_img, disps[0], init_disp, outputs["fwd"], outputs["inv"])
        loss_log.append(loss.detach().reshape(()))
        if epoch != opts["epochs"] - 1:
            loss.sum().backward()
            optim.step()
        out = {"disparity": disps[0].detach(), "poses": [p.detach() for p in poses],
               "poses_inv": [p.detach() for p in poses_inv]}
    out["losses"] = torch.stack(loss_log)
    return out
