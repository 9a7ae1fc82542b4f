import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_torch as O
from tcsfm_b200 import stn, synth
DEV = "cuda:0"
for seed in range(40):
    gen = torch.Generator().manual_seed(1234 + seed)
    b = int(torch.randint(1, 5, (1,), generator=gen)); h = int(torch.randint(18, 200, (1,), generator=gen)); w = int(torch.randint(18, 300, (1,), generator=gen))
    fr = synth.make_frames(b, h, w, seed=100 + seed, yaw=0.02, device=DEV, intrinsics=synth.scaled_intrinsics(h, w))
    K = fr["K"].clone()
    K[:, 0, 0] *= float(0.6 + 0.8 * torch.rand(1, generator=gen)); K[:, 1, 1] *= float(0.6 + 0.8 * torch.rand(1, generator=gen))
    K[:, 0, 1] = float(2.0 * torch.rand(1, generator=gen))
    scale = torch.tensor([0.05, 0.05, 0.3, 0.05, 0.18, 0.05])
    pose = (torch.randn(b, 6, generator=gen) * scale).to(DEV)
    if seed % 3 == 0: pose[:, 2] = -3.0
    depth_scale = float(0.2 + 3.0 * torch.rand(1, generator=gen))
    args = (fr["sources"][0], fr["depths"][0] * depth_scale, fr["depths"][1] * depth_scale, pose, K)
    ref, got = O.inverse_warp2(*args), stn.inverse_warp2(*args)
    bad = [(i, (got[i] != ref[i]).nonzero().tolist()[:6]) for i in range(4) if not torch.equal(got[i], ref[i])]
    if bad:
        print("seed", seed, "b,h,w", b, h, w, "HW", h * w, "HW%4", (h * w) % 4, bad)
        cam = O.backproject((fr["depths"][0] * depth_scale).squeeze(1), K.inverse())
        idx = bad[-1][1][0]
        print("   ref", ref[3][tuple(idx)].item(), "got", got[3][tuple(idx)].item())
print("done")
