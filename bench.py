#!/usr/bin/env python
"""Benchmark of the warp + SSIM/L1 photometric-loss hot path (BASELINE.json metric:
"warp+SSIM/L1 loss fwd+bwd frames/s at 192x640").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload kitti|scannet]

One *step* = one `Compute_Loss.forward` + `.backward()` over one synthetic minibatch
(config 2 of BASELINE.json: B=8 KITTI-shaped 3-frame snippets at 192x640, forward and
inverse direction for both sources = 32 pair evaluations, SSIM + L1 + auto-mask +
depth-consistency mask/term, gradients to the three disparity maps and four poses).
A *frame* is one target frame (2*S pair evaluations).

Prints ONE JSON line (rank 0).  Multi-GPU (torchrun, one rank per GPU): every rank
processes its own shard of minibatches, no data-path collective,
`scaling: weak`; the timed region is bracketed by barrier + synchronize and the
slowest rank's device time is used.

`--impl reference` times the oracle port of the reference's PyTorch CPU path
(oracle/ref_torch.py: the same ATen operators in the same order) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout must carry exactly one JSON line, but NCCL writes its version banner to fd 1: park the
# real stdout and route everything else that lands on fd 1 to stderr
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line_dict):
    os.write(_REAL_STDOUT, (json.dumps(line_dict) + "\n").encode())

PFT_WORKLOADS = {
    # sequential per-frame test-time optimisation (configs 3/4 of BASELINE.json); stand-in networks
    "pft": dict(b=6, h=192, w=640, n_src=2, iterations=4, epochs=20,
                desc="pft_kitti_windows_b6_192x640_4iter_20epochs_standin_nets"),
    "pft-scannet": dict(b=16, h=256, w=320, n_src=1, iterations=4, epochs=20,
                        desc="pft_scannet_pairs_b16_256x320_4iter_20epochs_standin_nets"),
}

WORKLOADS = {
    # name: (B, H, W, n_src, depth_range key, intrinsics)
    "kitti": dict(b=8, h=192, w=640, n_src=2, desc="kitti_triplets_b8_192x640_fwd+inv_ssim_l1_automask_depthconsist"),
    "scannet": dict(b=16, h=256, w=320, n_src=1, desc="scannet_pairs_b16_256x320_fwd+inv_ssim_l1_automask_depthconsist"),
    "scannet448": dict(b=16, h=256, w=448, n_src=1, desc="scannet_pairs_b16_256x448_fwd+inv_ssim_l1_automask_depthconsist"),
    # config 5: the training-step loss path at 4 scales (every scale nearest-upsampled to full
    # resolution, losses.py:86-87) at 376x1242
    "kitti376x4": dict(b=2, h=376, w=1242, n_src=2, scales=4,
                       desc="kitti_triplets_b2_376x1242_4scales_fwd+inv_ssim_l1_automask_depthconsist"),
}
LOSS_CFG = {"l1_weight": 0.15, "l_ssim_weight": 0.85, "l_smooth_weight": 0.05, "num_scales": 1,
            "l_depth_consist_weight": 0.14, "min_depth": 0.06, "max_depth": 2.67, "l_smooth": False,
            "l_reconstruction": True, "l_inverse": True, "l_depth_consist": True,
            "with_auto_mask": True, "l_ssim": True, "with_depth_mask": True}
METRIC = "warp+SSIM/L1 loss fwd+bwd frames/s at 192x640"
N_INPUT_SETS = 8      # rotating input sets: 8 x ~47 MB > 126 MB of L2, so no step finds its inputs in L2


def make_inputs(wl, seed, device, pin=False):
    from tcsfm_b200 import synth
    kitti = wl["h"] in (192, 376)
    rng = synth.KITTI_DEPTH_RANGE if kitti else synth.SCANNET_DEPTH_RANGE
    if wl["h"] == 376:
        base = torch.tensor(synth.KITTI_FULL_K, dtype=torch.float32)
    elif kitti:
        base = torch.tensor(synth.KITTI_K, dtype=torch.float32)
    else:
        base = synth.scaled_intrinsics(wl["h"], wl["w"], synth.SCANNET_K, (256, 320))
    fr = synth.make_frames(wl["b"], wl["h"], wl["w"], n_src=wl["n_src"], seed=seed, depth_range=rng, intrinsics=base)
    flat = {"target": fr["target"], "K": fr["K"]}
    for j in range(wl["n_src"]):
        flat["source%d" % j] = fr["sources"][j]
        flat["pose%d" % j] = fr["poses"][j]
        flat["pose_inv%d" % j] = fr["poses_inv"][j]
    for j in range(1 + wl["n_src"]):
        flat["disp%d" % j] = fr["disps"][j]
        for sc in range(1, wl.get("scales", 1)):          # lower-resolution disparities of the other scales
            flat["disp%d_s%d" % (j, sc)] = torch.nn.functional.avg_pool2d(fr["disps"][j], 2 ** sc, ceil_mode=True)
    if pin:
        return {k: v.pin_memory() for k, v in flat.items()}
    return {k: v.to(device) for k, v in flat.items()}


def run_step(loss_mod, inp, n_src, need_value=False):
    n_scales = loss_mod.num_scales
    disps = [inp["disp%d" % j].requires_grad_(True) for j in range(1 + n_src)]
    extra = [[inp["disp%d_s%d" % (j, sc)].requires_grad_(True) for sc in range(1, n_scales)] for j in range(1 + n_src)]
    poses = [inp["pose%d" % j].requires_grad_(True) for j in range(n_src)]
    poses_inv = [inp["pose_inv%d" % j].requires_grad_(True) for j in range(n_src)]
    for t in disps + poses + poses_inv + [e for ex in extra for e in ex]:
        t.grad = None
    out = loss_mod([inp["source%d" % j] for j in range(n_src)], inp["target"], [poses, poses_inv],
                   [[d] + ex for d, ex in zip(disps, extra)], inp["K"])
    total = out["total"]            # [1]; backward() on it directly, like train_mono.py:193
    total.backward()
    return total


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in o.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples if len(s) >= 6 for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_port_throughput(wl, steps, warmup, threads):
    """Times the oracle port of the reference's CPU path (fwd + bwd) on the host cores."""
    from oracle import ref_torch as O
    torch.set_num_threads(threads)
    inp = make_inputs(wl, 0, "cpu")
    n_src = wl["n_src"]

    n_scales = wl.get("scales", 1)
    kitti = wl["h"] in (192, 376)
    from tcsfm_b200 import synth
    rng = synth.KITTI_DEPTH_RANGE if kitti else synth.SCANNET_DEPTH_RANGE
    cfg = dict(LOSS_CFG, num_scales=n_scales, min_depth=rng[0], max_depth=rng[1])

    def step():
        disps = [[inp["disp%d" % j].clone().requires_grad_(True)] +
                 [inp["disp%d_s%d" % (j, sc)].clone().requires_grad_(True) for sc in range(1, n_scales)]
                 for j in range(1 + n_src)]
        poses = [inp["pose%d" % j].clone().requires_grad_(True) for j in range(n_src)]
        poses_inv = [inp["pose_inv%d" % j].clone().requires_grad_(True) for j in range(n_src)]
        out = O.compute_loss(cfg, [inp["source%d" % j] for j in range(n_src)], inp["target"],
                             [poses, poses_inv], disps, inp["K"])
        out["total"].sum().backward()
        return float(out["total"].sum().detach())

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return wl["b"] * steps / dt, dt / steps * 1e3


def main_pft(args):
    """PFT frames/s: every rank optimises its own contiguous shard of window minibatches
    (tcsfm_b200.shard), no communication; a 'step' is one window minibatch taken through all
    optimisation epochs (depth-net -> solve_pose_iteratively -> loss -> backward -> Adam)."""
    wl = PFT_WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from tcsfm_b200 import pft_driver, shard, synth
    opts = {"epochs": wl["epochs"], "num_source_imgs": wl["n_src"]}
    rng = synth.KITTI_DEPTH_RANGE if wl["h"] == 192 else synth.SCANNET_DEPTH_RANGE
    base = synth.KITTI_K if wl["h"] == 192 else synth.SCANNET_K
    steps = min(args.steps, 4)
    warm = max(1, min(args.warmup, 1))

    def frames_for(seed, device):
        return synth.make_frames(wl["b"], wl["h"], wl["w"], n_src=wl["n_src"], seed=seed, depth_range=rng,
                                 intrinsics=torch.tensor(base, dtype=torch.float32), device=device)

    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import ref_torch as O

        class OracleBackend:
            solve_pose_iteratively = staticmethod(O.iterative_pose)
            compute_optimization_loss = staticmethod(O.pft_window_loss)
            disp_to_depth = staticmethod(O.disp_to_depth)
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        cpu_opts = dict(opts, epochs=2)
        fr = frames_for(0, "cpu")
        dn, pn = synth.TinyDepthNet(0), synth.TinyPoseNet(0)
        t0 = time.perf_counter()
        pft_driver.optimize_window(dn, pn, fr["target"], fr["sources"], fr["K"], cpu_opts, wl["iterations"], rng, OracleBackend)
        dt = (time.perf_counter() - t0) * wl["epochs"] / cpu_opts["epochs"]
        fps = wl["b"] / dt
        emit(({"impl": "reference", "metric": "PFT frames/s", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
                          "steps": 1, "warmup": 0, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": wl["desc"]},
                          "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                           "sample": "2 of %d epochs of one window minibatch, extrapolated" % wl["epochs"]},
                          "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from tcsfm_b200 import _timing
    depth_net, pose_net = synth.TinyDepthNet(0).to(dev), synth.TinyPoseNet(0).to(dev)
    total_mbs = steps * world
    lo, hi = shard.shard_range(total_mbs, rank, world)           # this rank's window minibatches
    data = [frames_for(1000 + i, dev) for i in range(lo, hi)]
    host = [{k: ([t.cpu().pin_memory() for t in v] if isinstance(v, list) else v.cpu().pin_memory()) for k, v in d.items()}
            for d in data[:2]]

    runner = None if args.no_graph else pft_driver.WindowRunner(depth_net, pose_net, opts, wl["iterations"], rng)

    def run(fr):
        if runner is not None:        # epoch graphs captured on the first (warm-up) window, replayed afterwards
            return runner(fr["target"], fr["sources"], fr["K"])
        return pft_driver.optimize_window(depth_net, pose_net, fr["target"], fr["sources"], fr["K"], opts,
                                          wl["iterations"], rng)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warm):
        run(data[i % len(data)])
    barrier()
    l0 = _timing.LAUNCH_COUNT
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for fr in data:
        run(fr)
    e1.record()
    barrier()
    launches = _timing.LAUNCH_COUNT - l0
    ms = e0.elapsed_time(e1)
    # end to end: window inputs from pinned host memory, the final loss read back
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for i in range(len(data)):
        h = host[i % len(host)]
        fr = {k: ([t.to(dev, non_blocking=True) for t in v] if isinstance(v, list) else v.to(dev, non_blocking=True))
              for k, v in h.items() if k in ("target", "sources", "K")}
        float(run(fr)["losses"][-1])
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    # device time of the library's own launches inside one window minibatch (the hot path proper)
    timer = _timing.KernelTimer()
    with _timing.record(timer):                 # eager: events recorded during graph capture cannot be timed
        pft_driver.optimize_window(depth_net, pose_net, data[0]["target"], data[0]["sources"], data[0]["K"], opts,
                                   wl["iterations"], rng, cuda_graph=False)
    ksum = timer.summary()
    hot_ms = sum(v["launches"] * v["avg_ms"] for v in ksum.values())
    if dist is not None:
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
        dist.destroy_process_group()
    if rank != 0:
        return
    frames = wl["b"] * total_mbs
    h2d = sum(t.numel() * 4 for t in [host[0]["target"], host[0]["K"]] + host[0]["sources"])
    emit(({"metric": "PFT frames/s", "value": frames / (ms / 1e3), "unit": "frames/s", "n_gpus": world,
                      "steps": steps, "warmup": warm, "ms_per_step": ms / steps, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": wl["desc"], "window_minibatches_per_gpu": steps, "parallelism": "shard%d" % world,
                                 "networks": "stand-in TinyDepthNet/TinyPoseNet (the reference nets are out of scope)",
                                 "launch": "eager" if args.no_graph else "epoch CUDA graphs captured on the warm-up window, replayed for every timed window"},
                      "e2e": {"value": frames / (ms_e2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
                              "d2h_bytes_per_step": 4},
                      "gpu_launches": launches,
                      "hot_path": {"ms_per_window_minibatch": hot_ms, "frames_per_s": wl["b"] / (hot_ms / 1e3),
                                   "note": "sum of the library launches' device time (CUDA events), networks/optimiser excluded",
                                   "kernels": ksum}}))


def main_sweep(args):
    """Config 1's forward-only part: the loss-surface sweeps of the demo (plot_loss_surface.py:11-87 via
    helpers.compute_photometric_error): batch-1 evaluations of warp + photometric error for a line of
    translation and a line of yaw perturbations.  A launch-latency-bound use of the same kernels."""
    from tcsfm_b200 import pft, synth
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.sweep_evals
    if args.impl == "reference":
        from oracle import ref_torch as O
        dev, fn = "cpu", O.photometric_error
        torch.set_num_threads(os.cpu_count() or 1)
        n = min(n, 20)
    else:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dev, fn = "cuda", pft.compute_photometric_error
    fr = synth.make_frames(1, 192, 640, n_src=2, seed=0, device=dev, intrinsics=torch.tensor(synth.KITTI_K))
    deltas = torch.linspace(-0.05, 0.05, n)

    def sweep():
        vals = []
        with torch.no_grad():
            for axis in (2, 4):                                 # forward translation, yaw
                for d in deltas:
                    pose = fr["poses"][0].clone()
                    pose[:, axis] += d
                    r = fn(fr["target"], fr["sources"][0], fr["depths"][0], fr["depths"][1], pose, fr["K"])
                    vals.append((r["diff_img"] * r["valid_mask"] * r["weight_mask"]).sum() / r["valid_mask"].sum())
        return torch.stack(vals)

    sweep()
    if dev == "cuda":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = sweep()
    host = out.cpu()
    dt = time.perf_counter() - t0
    line = {"metric": "loss-surface evaluations/s (B=1, 192x640, forward only)", "value": 2 * n / dt, "unit": "evals/s",
            "n_gpus": 1, "steps": 2 * n, "warmup": 2 * n, "ms_per_step": dt / (2 * n) * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "demo_loss_surface_sweep_b1_192x640", "launch": "eager, wall clock incl. the final D2H"},
            "e2e": {"value": 2 * n / dt, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4}}
    if args.impl == "reference":
        line["impl"] = "reference"
        line["cpu_baseline"] = {"value": line["value"], "unit": "evals/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": "%d evaluations" % (2 * n)}
    else:
        line["surface_min"] = float(host.min())
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="kitti", choices=sorted(WORKLOADS) + sorted(PFT_WORKLOADS) + ["sweep"])
    ap.add_argument("--cpu-steps", type=int, default=None, help="steps of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of replaying CUDA graphs")
    ap.add_argument("--ddp-allreduce-mb", type=float, default=0.0,
                    help="multi-GPU only: after every step all-reduce a buffer of this many MB over NCCL, standing in "
                         "for DDP's exchange of the network gradients (62.9 MB for the reference's nets, SURVEY.md §5); "
                         "the loss path itself has no collective.  Implies eager launches.")
    ap.add_argument("--sweep-evals", type=int, default=100, help="workload 'sweep': evaluations per surface")
    ap.add_argument("--profile", default="full", choices=["full", "train"],
                    help="loss flags: 'full' = depth-consistency mask + term on (what PFT uses, 84 B/px/pair), "
                         "'train' = the paper's training defaults (run_mono_training.py:50-64: both off)")
    args = ap.parse_args()
    if args.workload in PFT_WORKLOADS:
        return main_pft(args)
    if args.workload == "sweep":
        return main_sweep(args)
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    config = {"workload": wl["desc"], "batch_per_gpu": wl["b"],
              "pairs_per_step_per_gpu": 2 * wl["n_src"] * wl["b"] * wl.get("scales", 1),
              "height": wl["h"], "width": wl["w"], "flags": "full (depth-consistency mask + term, auto-mask, SSIM+L1)",
              "parallelism": "shard%d" % max(world, args.gpus),
              "l2": "inputs rotate over %d sets (> L2 capacity)" % N_INPUT_SETS}

    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(args.steps, args.cpu_steps or 6))
        warm = max(1, min(args.warmup, 1))
        fps, ms = cpu_port_throughput(wl, steps, warm, cores)
        line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                 "sample": "%d steps of the same B=%d minibatch, oracle port of the reference's "
                                           "PyTorch CPU path, torch threads=%d" % (steps, wl["b"], cores)},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from tcsfm_b200 import _timing, losses, synth
    kitti = wl["h"] in (192, 376)
    rng = synth.KITTI_DEPTH_RANGE if kitti else synth.SCANNET_DEPTH_RANGE
    flags_cfg = {} if args.profile == "full" else {"l_depth_consist": False, "with_depth_mask": False}
    if args.profile == "train":
        config["flags"] = "paper training defaults (auto-mask, SSIM+L1; no depth-consistency mask/term)"
    loss_mod = losses.Compute_Loss(dict(LOSS_CFG, num_scales=wl.get("scales", 1), min_depth=rng[0], max_depth=rng[1],
                                        **flags_cfg))
    n_src = wl["n_src"]
    sets = [make_inputs(wl, 100 * rank + s, dev) for s in range(N_INPUT_SETS)]
    host_sets = [make_inputs(wl, 100 * rank + s, dev, pin=True) for s in range(2)]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    # The resident-input arm replays one captured CUDA graph per input set (the step has no
    # host-side control flow: the mean-on-mask threshold is decided on the device), so the
    # timed region contains exactly the device work of K full steps.
    graphs = None
    grad_buf = None
    if args.ddp_allreduce_mb > 0 and dist is not None:
        grad_buf = torch.zeros(int(args.ddp_allreduce_mb * 1e6 / 4), device=dev)
        args.no_graph = True
        config["collective"] = "NCCL all-reduce of %.1f MB per step (DDP stand-in)" % args.ddp_allreduce_mb
    if not args.no_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for s_ in sets:
                for _ in range(3):
                    run_step(loss_mod, s_, n_src)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graphs = []
        for s_ in sets:
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_):
                out_ = run_step(loss_mod, s_, n_src)
            graphs.append((g_, out_))
        config["launch"] = "cuda-graph replay of the full step (one graph per input set)"
    else:
        config["launch"] = "eager"

    def step_resident(i):
        if graphs is not None:
            graphs[i % N_INPUT_SETS][0].replay()
        else:
            run_step(loss_mod, sets[i % N_INPUT_SETS], n_src)
            if grad_buf is not None:
                dist.all_reduce(grad_buf)

    # End-to-end arm: every step copies its inputs from pinned host memory and reads the loss
    # back to the host.  Two device-side input buffers are used so that the copy of step i+1
    # (copy stream) overlaps the compute of step i (graph replay on the main stream); the loss
    # of step i is read on the host while step i+1 runs.
    loss_holder = [0.0]
    e2e = None
    if graphs is not None:
        copy_stream = torch.cuda.Stream()
        e2e = {"in": [graphs[0], graphs[1]], "bufs": [sets[0], sets[1]],
               "h2d_done": [torch.cuda.Event(), torch.cuda.Event()],
               "compute_done": [torch.cuda.Event(), torch.cuda.Event()],
               "loss_host": [torch.zeros(1, pin_memory=True), torch.zeros(1, pin_memory=True)]}
        for ev in e2e["compute_done"]:
            ev.record()

    def step_e2e(i):
        h = host_sets[i % 2]
        if e2e is None:
            inp = {k: v.to(dev, non_blocking=True) for k, v in h.items()}
            loss_holder[0] = float(run_step(loss_mod, inp, n_src).detach())   # device -> host read of the loss
            return
        k = i % 2
        with torch.cuda.stream(copy_stream), torch.no_grad():
            copy_stream.wait_event(e2e["compute_done"][k])          # buffer k is free again
            for name, src in h.items():
                e2e["bufs"][k][name].copy_(src, non_blocking=True)
            e2e["h2d_done"][k].record(copy_stream)
        main = torch.cuda.current_stream()
        main.wait_event(e2e["h2d_done"][k])
        graph_k, loss_k = e2e["in"][k]
        graph_k.replay()
        e2e["loss_host"][k].copy_(loss_k.detach(), non_blocking=True)
        e2e["compute_done"][k].record(main)
        if i > 0:                                                     # read the previous step's loss
            e2e["compute_done"][1 - k].synchronize()
            loss_holder[0] = float(e2e["loss_host"][1 - k])

    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = _timing.LAUNCH_COUNT
    ms_total = timed(step_resident, args.steps)
    launches = _timing.LAUNCH_COUNT - launches0
    clocks = sampler.summary() if sampler else None

    # per-launch device time of the library calls, live over a second (eager) pass of the same steps
    def step_eager(i):
        run_step(loss_mod, sets[i % N_INPUT_SETS], n_src)
    l0 = _timing.LAUNCH_COUNT
    timer = _timing.KernelTimer()
    with _timing.record(timer):
        timed(step_eager, min(args.steps, 50))
    ksum = timer.summary()
    if graphs is not None:       # launches inside a replayed graph are not seen by the Python counter
        launches = (_timing.LAUNCH_COUNT - l0) // min(args.steps, 50) * args.steps

    for i in range(min(args.warmup, 5)):
        step_e2e(i)
    e2e_steps = max(5, min(args.steps, 500))
    ms_e2e = timed(step_e2e, e2e_steps)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    frames = wl["b"] * world
    value = frames * args.steps / (ms_total / 1e3)
    e2e_value = frames * e2e_steps / (ms_e2e / 1e3)
    h2d = sum(v.numel() * v.element_size() for v in host_sets[0].values())
    npx = wl["h"] * wl["w"]
    pairs = 2 * n_src * wl["b"] * wl.get("scales", 1)
    # algorithmic bytes per pixel per pair (SURVEY.md §8d / DESIGN.md): fwd 32 R + 8 W, bwd 36 R + 8 W
    per_launch_pairs = 2 * n_src * wl["b"]
    bytes_per_launch = {"pair_loss_fwd": 40 * npx * per_launch_pairs, "pair_loss_bwd": 44 * npx * per_launch_pairs}
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    dom = max((k for k in ksum if k in bytes_per_launch), key=lambda k: ksum[k]["avg_ms"], default=None)
    roofline = None
    if dom:
        ach = bytes_per_launch[dom] / (ksum[dom]["avg_ms"] * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.isfile(tpath) and args.workload == "kitti":
            traffic = json.load(open(tpath)).get(dom)          # dram read+write per launch from the committed ncu capture
        roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "peak_source": peak_src, "avg_launch_ms": ksum[dom]["avg_ms"],
                    "algorithmic_bytes_per_launch": bytes_per_launch[dom],
                    "step_algorithmic_GBps": 84 * npx * pairs / (ms_total / args.steps * 1e-3) / 1e9,
                    "kernels": ksum}
    cpu_baseline = None
    if not args.no_cpu_baseline:
        csteps = args.cpu_steps or 4
        cfps, cms = cpu_port_throughput(wl, csteps, 1, cores)
        cpu_baseline = {"value": cfps, "unit": "frames/s", "cores": cores, "kind": "port", "ms_per_step": cms,
                        "sample": "%d steps of one B=%d minibatch of the same workload (oracle port of the "
                                  "reference's PyTorch CPU path, torch threads=%d)" % (csteps, wl["b"], cores)}
    metric = METRIC if args.workload == "kitti" else "warp+SSIM/L1 loss fwd+bwd frames/s at %dx%d" % (wl["h"], wl["w"])
    line = {"metric": metric, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "pairs_per_s": value * 2 * n_src, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
