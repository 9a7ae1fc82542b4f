"""Seeded synthetic inputs in the shapes the reference's loaders produce
(SURVEY.md §8d "Synthetic inputs").  There are no datasets in this image; every
test and benchmark says ``data: synthetic``.

Images are a low-frequency field plus fine texture; source frames are the target
shifted a few pixels horizontally plus noise (so the auto-mask is neither empty
nor full); disparities are sigmoids of smooth noise mapped to depth exactly like
``utils/learning_helpers.py:77-86``; poses are ~1 m/30 of forward motion with a
small yaw, the inverse pose is minus the forward one plus noise.
"""
import torch
import torch.nn.functional as F

KITTI_K = ((370.7, 0.0, 313.1), (0.0, 367.1, 94.6), (0.0, 0.0, 1.0))      # 192x640
SCANNET_K = ((288.8, 0.0, 159.9), (0.0, 308.7, 127.9), (0.0, 0.0, 1.0))   # 256x320
KITTI_FULL_K = ((718.856, 0.0, 607.19), (0.0, 718.856, 185.2), (0.0, 0.0, 1.0))  # 376x1242
KITTI_DEPTH_RANGE = (0.06, 2.67)     # run_mono_exps_kitti.sh:5
SCANNET_DEPTH_RANGE = (0.03, 3.0)    # run_scannet_exps.sh:3


def scaled_intrinsics(h, w, base=KITTI_K, base_hw=(192, 640)):
    k = torch.tensor(base, dtype=torch.float32).clone()
    k[0] *= w / base_hw[1]
    k[1] *= h / base_hw[0]
    return k


def _smooth(gen, b, c, h, w, cells=16):
    lo = torch.rand(b, c, max(2, h // cells + 1), max(2, w // cells + 1), generator=gen)
    return F.interpolate(lo, size=(h, w), mode="bilinear", align_corners=True)


def make_frames(b, h, w, n_src=2, seed=0, depth_range=KITTI_DEPTH_RANGE, intrinsics=None,
                yaw=0.01, device="cpu"):
    """Returns a dict: target [B,3,H,W], sources list of S x [B,3,H,W],
    disps [target, src...] each [B,1,H,W] in (0,1), depths (same order),
    poses / poses_inv lists of S x [B,6], K [B,3,3]."""
    gen = torch.Generator().manual_seed(seed)
    target = (_smooth(gen, b, 3, h, w) + 0.05 * torch.randn(b, 3, h, w, generator=gen)).clamp(0, 1)
    sources = []
    for j in range(n_src):
        shift = (3 + j) * (1 if j % 2 == 0 else -1)
        s = torch.roll(target, shifts=shift, dims=3) + 0.02 * torch.randn(b, 3, h, w, generator=gen)
        sources.append(s.clamp(0, 1).contiguous())
    disps, depths = [], []
    lo, hi = depth_range
    for _ in range(1 + n_src):
        d = torch.sigmoid(4.0 * (_smooth(gen, b, 1, h, w) - 0.5))
        disps.append(d)
        depths.append(1.0 / (1.0 / hi + (1.0 / lo - 1.0 / hi) * d))
    poses, poses_inv = [], []
    for j in range(n_src):
        sign = -1.0 if j % 2 == 0 else 1.0
        base = torch.tensor([0.001, 0.0005, sign * 0.03, 0.001, sign * yaw, 0.0005])
        p = base + 0.002 * torch.randn(b, 6, generator=gen)
        poses.append(p)
        poses_inv.append(-p + 0.0005 * torch.randn(b, 6, generator=gen))
    if intrinsics is None:
        intrinsics = scaled_intrinsics(h, w)
    k = intrinsics.unsqueeze(0).repeat(b, 1, 1).contiguous()
    out = {"target": target, "sources": sources, "disps": disps, "depths": depths,
           "poses": poses, "poses_inv": poses_inv, "K": k}

    def mv(x):
        if isinstance(x, list):
            return [mv(v) for v in x]
        return x.to(device)
    return {key: mv(val) for key, val in out.items()}


class TinyPoseNet(torch.nn.Module):
    """Deterministic stand-in for the reference pose network (models/pose_models.py
    is out of scope): 6-channel image stack [N,6,H,W] -> [N,6] pose, smooth in its
    input so gradients flow through the reconstructed image like in the real loop."""

    def __init__(self, seed=0, scale=0.01):
        super().__init__()
        gen = torch.Generator().manual_seed(seed)
        self.mix = torch.nn.Parameter(0.5 * torch.randn(6, 6, generator=gen))
        self.register_buffer("base", torch.tensor([0.001, 0.0005, -0.03, 0.001, -0.01, 0.0005]))
        self.scale = scale

    def forward(self, imgs):
        feat = imgs.mean(dim=(2, 3))                      # [N,6]
        sign = torch.sign(feat[:, 0:1] - feat[:, 3:4] + 1e-6).detach()
        return self.scale * torch.tanh(feat @ self.mix) + sign * 0 + self.base


class TinyDepthNet(torch.nn.Module):
    """Stand-in for the reference depth network (models/depth_models.py, out of scope):
    `encoder` + `decoder` sub-modules like the reference's, input [N,3,H,W] -> list with the
    scale-0 sigmoid disparity [N,1,H,W]."""

    def __init__(self, seed=0, width=8):
        super().__init__()
        torch.manual_seed(seed)
        self.encoder = torch.nn.Sequential(torch.nn.Conv2d(3, width, 3, padding=1), torch.nn.ELU(),
                                           torch.nn.Conv2d(width, width, 3, padding=1), torch.nn.ELU())
        self.decoder = torch.nn.Conv2d(width, 1, 3, padding=1)

    def forward(self, imgs):
        return [torch.sigmoid(self.decoder(self.encoder(imgs)))]
