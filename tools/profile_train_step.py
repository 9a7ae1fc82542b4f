"""Minimal driver for ncu: a few eager training steps of config 5 (stand-in networks), nothing else.
Usage: python tools/profile_train_step.py [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from tcsfm_b200 import ops, synth, training  # noqa: E402

ops.set_arithmetic(os.environ.get("TCSFM_ARITH", "fast"))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
wl = bench.TRAIN_WORKLOADS["train376x4"]
dev = torch.device("cuda:0")
cfg = training.default_config(num_scales=wl["scales"], iterations=wl["iterations"], full_profile=wl["full"])
step, optim = training.make_step(cfg, seed=0, device=dev, padded=True)
k = torch.tensor(synth.KITTI_FULL_K, dtype=torch.float32)
data = [synth.make_frames(wl["b"], wl["h"], wl["w"], n_src=2, seed=300 + i, intrinsics=k, device=dev) for i in range(2)]
for i in range(steps):
    training.run_train_step(step, optim, data[i % 2])
torch.cuda.synchronize()
print("done", steps)
