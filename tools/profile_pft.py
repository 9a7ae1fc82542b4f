"""Minimal driver for ncu: one PFT window minibatch (B=6, 192x640) for a few eager epochs."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tcsfm_b200 import pft_driver, synth
dev = torch.device("cuda:0")
epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 3
fr = synth.make_frames(6, 192, 640, n_src=2, seed=0, device=dev, intrinsics=torch.tensor(synth.KITTI_K))
dn, pn = synth.TinyDepthNet(0).to(dev), synth.TinyPoseNet(0).to(dev)
out = pft_driver.optimize_window(dn, pn, fr["target"], fr["sources"], fr["K"], {"epochs": epochs}, 4)
torch.cuda.synchronize()
print("losses", out["losses"].tolist())
