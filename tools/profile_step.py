"""Minimal driver for ncu: W warm-up + K steps of the bench workload (config 2), nothing else."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from tcsfm_b200 import losses, ops  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wl = bench.WORKLOADS[sys.argv[3] if len(sys.argv) > 3 else "kitti"]
ops.set_arithmetic(os.environ.get("TCSFM_ARITH", "exact"))
dev = torch.device("cuda:0")
mod = losses.Compute_Loss(bench.LOSS_CFG)
sets = [bench.make_inputs(wl, s, dev) for s in range(2)]
for i in range(warm + steps):
    bench.run_step(mod, sets[i % 2], wl["n_src"])
torch.cuda.synchronize()
print("profiled %d steps" % steps)
