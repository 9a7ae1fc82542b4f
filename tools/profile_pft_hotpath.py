"""Driver for ncu / event timing of the PFT hot path proper: per optimisation epoch the reference runs
disp_to_depth -> solve_pose_iteratively(4 iterations, return_errors=True) -> compute_optimization_loss -> backward
(optimization_experiments/optimizer.py:241-268).  The depth network is replaced by leaf disparity tensors (it is out of
scope and dominates a window with any stand-in), the pose network is the cheap stand-in; what remains besides the
library's kernels is the PyTorch glue between them.  Usage: python tools/profile_pft_hotpath.py [epochs] [B H W]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tcsfm_b200 import _timing, losses, pft, pft_driver, synth, train_mono  # noqa: E402


def run(epochs=3, b=6, h=192, w=640, n_src=2, iterations=4, dev="cuda:0", record=False):
    fr = synth.make_frames(b, h, w, n_src=n_src, seed=0, device=dev, intrinsics=synth.scaled_intrinsics(h, w))
    pose_net = synth.TinyPoseNet(0).to(dev)
    disps = [d.clone().requires_grad_(True) for d in fr["disps"]]
    opts = dict(pft_driver.DEFAULT_OPTIONS, num_source_imgs=n_src)
    init = fr["disps"][0] * 0.9 + 0.02

    def epoch():
        for d in disps:
            d.grad = None
        depths = [losses.disp_to_depth(d, *synth.KITTI_DEPTH_RANGE)[1] for d in disps]
        _, _, out = train_mono.solve_pose_iteratively(iterations, depths, pose_net, fr["target"], fr["sources"], fr["K"],
                                                      return_errors=True)
        loss = pft.compute_optimization_loss(opts, fr["target"], disps[0], init, out["fwd"], out["inv"])
        loss.sum().backward()
        return loss
    for _ in range(2):
        epoch()
    torch.cuda.synchronize()
    timer = _timing.KernelTimer()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if record:
        # whole-epoch device time from a CUDA-graph replay (no launch gaps), library time from eager events
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            epoch()
        torch.cuda.current_stream().wait_stream(side)
        with torch.cuda.graph(g):
            epoch()
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(epochs):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        epoch_ms = e0.elapsed_time(e1) / epochs
        with _timing.record(timer):
            for _ in range(epochs):
                # the device is parked on a ~10 ms spin so that the whole epoch is queued before it starts:
                # the event pairs then bracket device time only, not the host's submission gaps
                torch.cuda._sleep(20_000_000)
                epoch()
                torch.cuda.synchronize()
        ksum = timer.summary()
        lib_ms = sum(v["launches"] * v["avg_ms"] for v in ksum.values()) / epochs
        return {"epoch_ms_graph_replay": epoch_ms, "library_ms_per_epoch": lib_ms, "library_share": lib_ms / epoch_ms,
                "kernels": ksum}
    for _ in range(epochs):
        loss = epoch()
    torch.cuda.synchronize()
    return float(loss.detach())


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    epochs = a[0] if a else 3
    shape = a[1:4] if len(a) >= 4 else [6, 192, 640]
    print("loss", run(epochs, *shape))
