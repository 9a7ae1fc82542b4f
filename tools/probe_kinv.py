"""Probe behind the own K^-1 kernel: torch.linalg.inv_ex (cuBLAS batched LU + triangular solves on identity) of
several families of 3x3 matrices on the GPU, saved with their inputs so that candidate orderings of the LU arithmetic
can be matched bit for bit offline (tools/match_kinv.py).  Usage: python tools/probe_kinv.py out.npz [n]"""
import sys

import numpy as np
import torch

out = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
g = torch.Generator().manual_seed(0)


def intrinsics(n, skew=False):
    fx = 200 + 1000 * torch.rand(n, generator=g)
    fy = fx * (0.9 + 0.2 * torch.rand(n, generator=g))
    cx = 100 + 600 * torch.rand(n, generator=g)
    cy = 50 + 300 * torch.rand(n, generator=g)
    k = torch.zeros(n, 3, 3)
    k[:, 0, 0], k[:, 1, 1], k[:, 0, 2], k[:, 1, 2], k[:, 2, 2] = fx, fy, cx, cy, 1.0
    if skew:
        k[:, 0, 1] = 5 * torch.randn(n, generator=g)
    return k


fam = {
    "kitti": intrinsics(n),
    "skew": intrinsics(n, skew=True),
    "dense": torch.randn(n, 3, 3, generator=g) + 3 * torch.eye(3),
    "general": torch.randn(n, 3, 3, generator=g) * torch.tensor([300.0, 30.0, 1.0]).view(1, 3, 1),
    "lower": intrinsics(n).transpose(1, 2).contiguous(),
}
res = {}
for name, k in fam.items():
    kd = k.cuda()
    inv, info = torch.linalg.inv_ex(kd)
    # batch 1 may take a different path: probe a few singly
    single = torch.stack([torch.linalg.inv_ex(kd[i:i + 1])[0][0] for i in range(8)])
    res[name + "_in"] = k.numpy()
    res[name + "_inv"] = inv.cpu().numpy()
    res[name + "_single"] = single.cpu().numpy()
    res[name + "_inverse_fn"] = torch.inverse(kd[:64]).cpu().numpy()
    print(name, "batch == single:", bool(torch.equal(inv[:8], single)), " inverse() == inv_ex:", bool(torch.equal(torch.inverse(kd[:64]), inv[:64])),
          "strides", inv.stride())
np.savez(out, **res)
print("saved", out)
