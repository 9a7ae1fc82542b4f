"""Scratch GPU diagnostic: per-gradient rel-L2 of the drop-in Compute_Loss vs the eager CUDA oracle."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import goldens
from goldens import rel_l2
from oracle import ref_torch as O
from tcsfm_b200 import losses, synth
DEV = "cuda:0"
def leaf(t): return t.clone().detach().requires_grad_(True)
SHAPES = [(4, 192, 640, 0.01, synth.KITTI_DEPTH_RANGE), (4, 256, 320, 0.04, synth.SCANNET_DEPTH_RANGE),
          (2, 256, 448, 0.02, synth.SCANNET_DEPTH_RANGE), (1, 376, 1242, 0.02, synth.KITTI_DEPTH_RANGE),
          (3, 50, 77, 0.08, synth.KITTI_DEPTH_RANGE), (4, 256, 320, 0.04, synth.KITTI_DEPTH_RANGE), (4, 256, 320, 0.01, synth.SCANNET_DEPTH_RANGE)]
for (b, h, w, yaw, rng) in SHAPES:
    fr = synth.make_frames(b, h, w, seed=21, yaw=yaw, depth_range=rng, device=DEV, intrinsics=synth.scaled_intrinsics(h, w))
    for tag in ("train", "full"):
        cfg = dict(goldens.LOSS_CFGS[tag], min_depth=rng[0], max_depth=rng[1])
        res = []
        for impl in ("oracle", "cuda", "oracle64"):
            dt = torch.float64 if impl == "oracle64" else torch.float32
            disps = [leaf(d.to(dt)) for d in fr["disps"]]
            poses, poses_inv = [leaf(p.to(dt)) for p in fr["poses"]], [leaf(p.to(dt)) for p in fr["poses_inv"]]
            dl = [[disps[0]], [disps[1]], [disps[2]]]
            srcs = [s.to(dt) for s in fr["sources"]]
            if impl != "cuda":
                out = O.compute_loss(cfg, srcs, fr["target"].to(dt), [poses, poses_inv], dl, fr["K"].to(dt))
            else:
                out = losses.Compute_Loss(cfg)(srcs, fr["target"], [poses, poses_inv], dl, fr["K"])
            out["total"].sum().backward()
            res.append((out, disps, poses, poses_inv))
        (ro, rd, rp, rpi), (go, gd, gp, gpi), (o6, d6, p6, pi6) = res
        z = lambda t, like: t.grad if t.grad is not None else torch.zeros_like(like)
        msg = ["%dx%dx%d yaw%.2f %s" % (b, h, w, yaw, tag)]
        for k in ("l_reconstruct_inverse", "l_reconstruct_forward", "l_depth", "total"):
            msg.append("%s %.3e/%.3e" % (k[:9], abs(float(go[k]) - float(ro[k])) / max(abs(float(ro[k])), 1e-12), abs(float(o6[k]) - float(ro[k])) / max(abs(float(ro[k])), 1e-12)))
        for j in range(3):
            msg.append("gdisp%d new-ref32 %.2e new-ref64 %.2e ref32-ref64 %.2e" % (j, rel_l2(z(gd[j], rd[j]), z(rd[j], rd[j])), rel_l2(z(gd[j], rd[j]), z(d6[j], rd[j])), rel_l2(z(rd[j], rd[j]), z(d6[j], rd[j]))))
        for j in range(2):
            msg.append("gpose%d %.2e/%.2e gposeinv%d %.2e/%.2e" % (j, rel_l2(z(gp[j], rp[j]), z(rp[j], rp[j])), rel_l2(z(rp[j], rp[j]), z(p6[j], rp[j])), j, rel_l2(z(gpi[j], rpi[j]), z(rpi[j], rpi[j])), rel_l2(z(rpi[j], rpi[j]), z(pi6[j], rpi[j]))))
        print("\n  ".join(msg))
