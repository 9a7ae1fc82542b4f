"""Training-step driver around the fused hot path, mirroring one minibatch of the reference's
``Trainer.forward`` (train_mono.py:159-194):

    disparities = solve_disp(depth_model, target, sources)                 # train_mono.py:166
    depths      = [disp_to_depth(d[0], min_depth, max_depth)[1] ...]       # :171
    poses, inv  = solve_pose_iteratively(iterations, depths, pose_model, target, sources, K)   # :173
    losses      = Compute_Loss(...)(sources, target, [poses, inv], disparities, K)              # :176
    total       = losses['total'] + l_pose_consist_weight * pose_consistency(poses, inv)        # :179-181
    total.backward(); optimizer.step()                                                          # :192-194

The reference is single-GPU.  Multi-GPU training is data parallel over the minibatch:
one process per GPU, every rank evaluates the loss on its own sub-batch (the masked means normalise by
the rank's own mask sums, exactly like a single-GPU run on that sub-batch) and the ONLY collective is
``DistributedDataParallel``'s NCCL all-reduce (mean) of the network gradients -- the fused loss kernels
themselves never communicate.

The depth / pose networks are out of scope (they stay the reference's PyTorch modules); the stand-ins
below only give the step something differentiable of the reference's parameter volume to drive:
14.1 M + 1.59 M fp32 parameters = 62.9 MB of gradients per step.
"""
import torch

from . import losses as _losses
from . import synth, train_mono

REFERENCE_DEPTH_PARAMS = 14_100_000      # models/depth_models.py (ResNet-18 encoder + decoder), SURVEY.md §2.1 row 8
REFERENCE_POSE_PARAMS = 1_590_000        # models/pose_models.py, row 9


class StandInDepthNet(torch.nn.Module):
    """[N,3,H,W] -> list of `num_scales` sigmoid disparities [N,1,H>>s,W>>s] (ceil sizes), like the
    reference decoder's multi-scale outputs.  `ballast` pads the parameter count to the reference's so
    that the gradient all-reduce moves the reference's 56 MB for this network."""

    def __init__(self, seed=0, num_scales=1, width=8, n_params=REFERENCE_DEPTH_PARAMS):
        super().__init__()
        self.core = synth.TinyDepthNet(seed, width)
        self.num_scales = num_scales
        own = sum(p.numel() for p in self.core.parameters())
        # (no zero-element parameter: DDP would wait for a gradient that never arrives)
        self.ballast = torch.nn.Parameter(torch.zeros(n_params - own)) if n_params > own else None

    @property
    def encoder(self):
        return self.core.encoder

    def forward(self, imgs):
        disp = self.core(imgs)[0]
        if self.ballast is not None:
            # every parameter takes part in the step (DDP reduces all of them): a vanishing weight-decay-like term
            disp = disp + 1e-12 * self.ballast.square().mean()
        out = [disp]
        for s in range(1, self.num_scales):
            out.append(torch.nn.functional.avg_pool2d(disp, 2 ** s, ceil_mode=True))
        return out


class StandInPoseNet(torch.nn.Module):
    def __init__(self, seed=0, n_params=REFERENCE_POSE_PARAMS):
        super().__init__()
        self.core = synth.TinyPoseNet(seed)
        own = sum(p.numel() for p in self.core.parameters())
        self.ballast = torch.nn.Parameter(torch.zeros(n_params - own)) if n_params > own else None

    def forward(self, imgs):
        pose = self.core(imgs)
        if self.ballast is not None:
            pose = pose + 1e-12 * self.ballast.square().mean()
        return pose


class Backend:
    """The hot-path callables of the step; tests swap in the oracle's."""
    solve_pose_iteratively = staticmethod(train_mono.solve_pose_iteratively)
    compute_pose_consistency_loss = staticmethod(train_mono.compute_pose_consistency_loss)
    disp_to_depth = staticmethod(_losses.disp_to_depth)

    @staticmethod
    def make_loss(config):
        return _losses.Compute_Loss(config)


class TrainStep(torch.nn.Module):
    """forward(target, sources..., K) -> total loss [1] of one minibatch (train_mono.py:166-181).  A module so
    that ``DistributedDataParallel(TrainStep(...))`` hooks the gradient all-reduce onto its backward."""

    def __init__(self, depth_net, pose_net, config, backend=Backend):
        super().__init__()
        self.depth_net, self.pose_net = depth_net, pose_net
        self.config = config
        self.backend = backend
        self.loss = backend.make_loss(config)
        self.last_losses = None

    def forward(self, target, *rest):
        sources, K = list(rest[:-1]), rest[-1]
        cfg = self.config
        n = target.shape[0]
        disp_all = self.depth_net(torch.cat([target] + sources, 0))                     # solve_disp, train_mono.py:122-132
        disparities = [[d[j * n:(j + 1) * n] for d in disp_all] for j in range(1 + len(sources))]
        depths = [self.backend.disp_to_depth(d[0], cfg['min_depth'], cfg['max_depth'])[1] for d in disparities]
        poses, poses_inv = self.backend.solve_pose_iteratively(cfg['iterations'], depths, self.pose_net, target, sources, K)
        out = self.loss(sources, target, [poses, poses_inv], disparities, K)
        total = out['total']
        if cfg.get('l_pose_consist', True):
            total = total + cfg.get('l_pose_consist_weight', 5) * self.backend.compute_pose_consistency_loss(poses, poses_inv)
        self.last_losses = out
        return total


def default_config(num_scales=1, iterations=4, depth_range=synth.KITTI_DEPTH_RANGE, full_profile=False):
    """run_mono_training.py:27-64 defaults (the paper's training flags); full_profile adds the
    depth-consistency mask + term that PFT uses."""
    return {"l1_weight": 0.15, "l_ssim_weight": 0.85, "l_smooth_weight": 0.05, "num_scales": num_scales,
            "l_depth_consist_weight": 0.14, "min_depth": depth_range[0], "max_depth": depth_range[1], "l_smooth": True,
            "l_reconstruction": True, "l_inverse": True, "l_depth_consist": bool(full_profile),
            "with_auto_mask": True, "l_ssim": True, "with_depth_mask": bool(full_profile),
            "l_pose_consist": True, "l_pose_consist_weight": 5, "iterations": iterations}


def make_step(config, seed=0, device="cpu", padded=True, backend=Backend, lr=9e-4, capturable=False):
    """(TrainStep, Adam) with seeded stand-in networks; `padded` = reference-sized parameter volume;
    `capturable`: Adam keeps its step counter on the device so that `optim.step()` can live in a CUDA graph."""
    depth = StandInDepthNet(seed, config['num_scales'], n_params=REFERENCE_DEPTH_PARAMS if padded else 0)
    pose = StandInPoseNet(seed, n_params=REFERENCE_POSE_PARAMS if padded else 0)
    step = TrainStep(depth, pose, config, backend).to(device)
    optim = torch.optim.Adam(step.parameters(), lr=lr, capturable=capturable)            # run_mono_training.py:155
    return step, optim


def wrap_ddp(step, device=None, bucket_cap_mb=25):
    """DistributedDataParallel around the step: NCCL (or gloo in the CPU tests) all-reduce of the network
    gradients, bucketed and overlapped with the remaining backward."""
    from torch.nn.parallel import DistributedDataParallel as DDP
    ids = [device.index] if device is not None and device.type == "cuda" else None
    return DDP(step, device_ids=ids, bucket_cap_mb=bucket_cap_mb, gradient_as_bucket_view=True)


def run_train_step(model, optim, frames, sync=True):
    """One optimisation step on `frames` (synth.make_frames layout).  sync=False skips DDP's all-reduce
    (`no_sync`), which is how the benchmark separates the collective's exposed time."""
    import contextlib
    optim.zero_grad(set_to_none=True)
    ctx = model.no_sync() if (not sync and hasattr(model, "no_sync")) else contextlib.nullcontext()
    with ctx:
        total = model(frames["target"], *frames["sources"], frames["K"])
        total.sum().backward()
    optim.step()
    return total.detach()


class FlatGradTrainer:
    """The same optimisation step with the host taken out of it: every parameter's gradient is a view of ONE flat
    buffer, forward + backward are captured as one CUDA graph and the Adam update as a second one; between the two
    replays the data-parallel form issues a single NCCL all-reduce (mean) of the flat buffer -- still the only
    collective of the mode, but one launch over the whole 62.8 MB instead of DDP's per-bucket calls interleaved with a
    host-launch-bound backward.  The eager step launches ~1 200 kernels from Python and is bound by that; the replayed
    step is bound by the device.

    `use_graph=False` runs the identical sequence eagerly (CPU / gloo tests, debugging).  The networks must not
    change shape between steps and the frames must keep the shapes of `example` (static input buffers)."""

    def __init__(self, step, optim, example, group=None, use_graph=True, warmup=3):
        self.step, self.optim, self.group = step, optim, group
        self.params = [p for p in step.parameters() if p.requires_grad]
        dev = self.params[0].device
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.static = {"target": example["target"].clone(), "sources": [t.clone() for t in example["sources"]],
                       "K": example["K"].clone()}
        self.total = None
        self.graph_fb = self.graph_opt = None
        if use_graph:
            if dev.type != "cuda":
                raise RuntimeError("FlatGradTrainer: CUDA graphs need CUDA parameters (use_graph=False for the eager form)")
            # warm-up on a side stream (cuDNN plans, lazily created Adam state), then put parameters and optimizer
            # state back IN PLACE -- the graphs hold their addresses -- so that building the trainer is not a step
            saved_p = [p.detach().clone() for p in self.params]
            saved_s = {id(t): t.clone() for st in self.optim.state.values() for t in st.values() if torch.is_tensor(t)}
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    self._forward_backward()
                    self._reduce()
                    self.optim.step()
            torch.cuda.current_stream().wait_stream(side)
            with torch.no_grad():
                for p, q in zip(self.params, saved_p):
                    p.copy_(q)
                for st in self.optim.state.values():
                    for t in st.values():
                        if torch.is_tensor(t):
                            t.copy_(saved_s[id(t)]) if id(t) in saved_s else t.zero_()
            torch.cuda.synchronize()
            self.graph_fb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_fb):
                self._forward_backward()
            self.graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_opt, pool=self.graph_fb.pool()):
                self.optim.step()

    def _forward_backward(self):
        self.flat.zero_()                       # the gradients accumulate in place into the views of `flat`
        st = self.static
        total = self.step(st["target"], *st["sources"], st["K"])
        total.sum().backward()
        self.total = total.detach()

    def _reduce(self, sync=True):
        if self.group is None or not sync:
            return
        import torch.distributed as dist
        world = dist.get_world_size(self.group)
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
        else:                                   # gloo has no AVG
            dist.all_reduce(self.flat, group=self.group)
            self.flat.div_(world)

    def run(self, frames, sync=True):
        """One optimisation step on `frames` (synth.make_frames layout); returns the total loss (a static tensor that
        the next step overwrites)."""
        with torch.no_grad():
            self.static["target"].copy_(frames["target"], non_blocking=True)
            for dst, src in zip(self.static["sources"], frames["sources"]):
                dst.copy_(src, non_blocking=True)
            self.static["K"].copy_(frames["K"], non_blocking=True)
        if self.graph_fb is not None:
            self.graph_fb.replay()
            self._reduce(sync)
            self.graph_opt.replay()
        else:
            self._forward_backward()
            self._reduce(sync)
            self.optim.step()
        return self.total
