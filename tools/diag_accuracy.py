"""Prints the actual loss / gradient agreement of Compute_Loss with the oracle run by eager PyTorch on the same GPU\n(configs 2 and 3).  The numbers quoted in README.md / DESIGN.md come from here."""
import sys, torch
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from tcsfm_b200 import losses, synth
from oracle import ref_torch as O
import goldens
dev = torch.device('cuda:0')
def leaf(t): return t.clone().detach().requires_grad_(True)
def rel(a, b): return float((a - b).norm() / b.norm().clamp_min(1e-30))
for (b, h, w, nsrc, rng, K, hw) in [(8, 192, 640, 2, synth.KITTI_DEPTH_RANGE, synth.KITTI_K, (192, 640)), (16, 256, 320, 1, synth.SCANNET_DEPTH_RANGE, synth.SCANNET_K, (256, 320))]:
    fr = synth.make_frames(b, h, w, n_src=nsrc, seed=0, depth_range=rng, device=dev, intrinsics=synth.scaled_intrinsics(h, w, K, hw))
    cfg = dict(goldens.LOSS_CFGS["full"], min_depth=rng[0], max_depth=rng[1])
    res = []
    for impl in ("oracle", "cuda"):
        disps = [[leaf(d)] for d in fr["disps"]]
        poses, poses_inv = [leaf(p) for p in fr["poses"]], [leaf(p) for p in fr["poses_inv"]]
        args = (fr["sources"], fr["target"], [poses, poses_inv], disps, fr["K"])
        out = O.compute_loss(cfg, *args) if impl == "oracle" else losses.Compute_Loss(cfg)(*args)
        out["total"].sum().backward()
        res.append((out, [d[0] for d in disps] + poses + poses_inv))
    (ro, rl), (go, gl) = res
    print((b, h, w), "loss rel", {k: abs(float(go[k]) - float(ro[k])) / abs(float(ro[k])) for k in ("l_reconstruct_inverse", "l_reconstruct_forward", "l_depth", "total")})
    print("   grad rel_l2", [round(rel(a.grad, bb.grad), 9) for a, bb in zip(gl, rl)])
