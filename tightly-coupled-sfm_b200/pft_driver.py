"""Synthetic-data driver for per-frame test-time optimisation (PFT), mirroring the loop of
the reference's `DepthOptimizer.optimize_window` (optimization_experiments/optimizer.py:
136-297) around the fused hot path: per window minibatch a private copy of the depth
network is optimised for `epochs` steps of  depth-net -> disp_to_depth ->
solve_pose_iteratively(return_errors=True) -> compute_optimization_loss -> backward -> Adam.

The reference's networks, loaders, ScaleRecovery, plotting and trajectory stitching are out
of scope; the stand-in networks (synth.TinyDepthNet / TinyPoseNet) only give the loop
something differentiable to drive.  Windows are independent, so a sequence is sharded across
ranks with `shard.shard_range` and no communication.
"""
import copy

import torch

from . import losses, pft, train_mono

DEFAULT_OPTIONS = {
    # optimization_experiments/run_sequential_optimization.py:69-99 (loss-relevant keys)
    "num_source_imgs": 2, "diff_img_argmin": True, "automasking": True, "l_inverse_reconstruction": True,
    "l_depth_consist": True, "l_depth_consist_weight": 0.15, "l_depth_init": True, "l_depth_init_weight": 0.1,
    "l_smooth": False, "l_smooth_weight": 0.05, "l_pose_consist": False, "plotting": False,
    "epochs": 20, "lr": 2e-4,
}


class Backend:
    """The three hot-path callables the loop needs; tests swap in the oracle's."""
    solve_pose_iteratively = staticmethod(train_mono.solve_pose_iteratively)
    compute_optimization_loss = staticmethod(pft.compute_optimization_loss)
    disp_to_depth = staticmethod(losses.disp_to_depth)


def optimize_window(depth_net, pose_net, target_img, source_imgs, intrinsics, options=None, iterations=4,
                    depth_range=(0.06, 2.67), backend=Backend, cuda_graph=False):
    """One window minibatch (optimizer.py:136-297).  Returns dict(losses=[per-epoch loss tensors],
    disparity=final target disparity, poses=..., poses_inv=...).

    cuda_graph=True captures one optimisation epoch (networks, hot path, backward, Adam) into a
    CUDA graph after three eager epochs and replays it for the rest: the fused path has no host
    synchronisation or data-dependent control flow, so the whole epoch is capturable."""
    opts = dict(DEFAULT_OPTIONS, **(options or {}))
    bsz = target_img.shape[0]
    imgs = torch.cat([target_img] + list(source_imgs), 0)
    with torch.no_grad():                                    # un-optimised prediction, optimizer.py:143-160
        init_disp = depth_net(imgs)[0][0:bsz].clone()
    net = copy.deepcopy(depth_net)                            # optimizer.py:177-182
    optim = torch.optim.Adam(net.encoder.parameters(), lr=opts["lr"], capturable=bool(cuda_graph))
    state = {}

    def forward():                                            # optimizer.py:217-263
        disp = net(imgs)[0]
        disps = [disp[i * bsz:(i + 1) * bsz] for i in range(1 + len(source_imgs))]
        depths = [backend.disp_to_depth(d, depth_range[0], depth_range[1])[1] for d in disps]
        poses, poses_inv, outputs = backend.solve_pose_iteratively(
            iterations, depths, pose_net, target_img, list(source_imgs), intrinsics, return_errors=True)
        loss = backend.compute_optimization_loss(opts, target_img, disps[0], init_disp, outputs["fwd"], outputs["inv"])
        state.update(disparity=disps[0].detach(), poses=[p.detach() for p in poses],
                     poses_inv=[p.detach() for p in poses_inv])
        return loss

    def train_step():                                         # optimizer.py:266-268
        optim.zero_grad(set_to_none=True)
        loss = forward()
        loss.sum().backward()
        optim.step()
        return loss

    loss_log = []
    n_train = opts["epochs"] - 1                              # the last epoch only evaluates
    if not cuda_graph:
        for _ in range(n_train):
            loss_log.append(train_step().detach().reshape(()))
    else:
        warm = min(3, n_train)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warm):
                loss_log.append(train_step().detach().reshape(()).clone())
        torch.cuda.current_stream().wait_stream(side)
        if n_train > warm:
            graph = torch.cuda.CUDAGraph()
            optim.zero_grad(set_to_none=True)
            with torch.cuda.graph(graph):
                static_loss = train_step()
            for _ in range(n_train - warm):
                graph.replay()
                loss_log.append(static_loss.detach().reshape(()).clone())
    with torch.no_grad():
        loss_log.append(forward().detach().reshape(()))
    out = dict(state)
    out["losses"] = torch.stack(loss_log)
    return out


class WindowRunner:
    """Optimises many window minibatches of the same shape (a sequence shard) and captures the
    optimisation epoch only once: the first window runs three eager epochs, captures one training
    epoch and one evaluation epoch as CUDA graphs, and every later window just resets the private
    network copy / Adam state, refreshes the static input buffers and replays the graphs.

    Equivalent to calling `optimize_window` per window (same kernels in the same order); valid
    because the fused path has no host synchronisation and no data-dependent control flow."""

    def __init__(self, depth_net, pose_net, options=None, iterations=4, depth_range=(0.06, 2.67), backend=Backend):
        self.opts = dict(DEFAULT_OPTIONS, **(options or {}))
        self.depth_net, self.pose_net = depth_net, pose_net
        self.iterations, self.depth_range, self.backend = iterations, depth_range, backend
        self.net = copy.deepcopy(depth_net)
        self.optim = torch.optim.Adam(self.net.encoder.parameters(), lr=self.opts["lr"], capturable=True)
        self.train_graph = self.eval_graph = None
        self.static = None

    def _reset(self):
        with torch.no_grad():
            for p, p0 in zip(self.net.parameters(), self.depth_net.parameters()):
                p.copy_(p0)
            for b, b0 in zip(self.net.buffers(), self.depth_net.buffers()):
                b.copy_(b0)
            for st in self.optim.state.values():
                for v in st.values():
                    if torch.is_tensor(v):
                        v.zero_()

    def _forward(self):
        s, bsz = self.static, self.static["target"].shape[0]
        disp = self.net(s["imgs"])[0]
        n_src = len(s["sources"])
        disps = [disp[i * bsz:(i + 1) * bsz] for i in range(1 + n_src)]
        depths = [self.backend.disp_to_depth(d, self.depth_range[0], self.depth_range[1])[1] for d in disps]
        poses, poses_inv, outputs = self.backend.solve_pose_iteratively(
            self.iterations, depths, self.pose_net, s["target"], list(s["sources"]), s["K"], return_errors=True)
        loss = self.backend.compute_optimization_loss(self.opts, s["target"], disps[0], s["init_disp"],
                                                      outputs["fwd"], outputs["inv"])
        return loss, disps[0], poses, poses_inv

    def _train_step(self):
        self.optim.zero_grad(set_to_none=True)
        loss = self._forward()[0]
        loss.sum().backward()
        self.optim.step()
        return loss

    def _set_inputs(self, target, sources, K):
        if self.static is None:
            self.static = {"target": target.clone(), "sources": [s.clone() for s in sources], "K": K.clone()}
            self.static["imgs"] = torch.cat([self.static["target"]] + self.static["sources"], 0)
            bsz = target.shape[0]
            with torch.no_grad():
                self.static["init_disp"] = self.depth_net(self.static["imgs"])[0][0:bsz].clone()
            return
        s = self.static
        with torch.no_grad():
            s["target"].copy_(target)
            for dst, src in zip(s["sources"], sources):
                dst.copy_(src)
            s["imgs"].copy_(torch.cat([s["target"]] + s["sources"], 0))
            s["K"].copy_(K)
            s["init_disp"].copy_(self.depth_net(s["imgs"])[0][0:target.shape[0]])

    def __call__(self, target_img, source_imgs, intrinsics):
        n_train = self.opts["epochs"] - 1
        self._set_inputs(target_img, source_imgs, intrinsics)
        self._reset()
        losses = []
        if self.train_graph is None:
            warm = min(3, n_train)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warm):
                    losses.append(self._train_step().detach().reshape(()).clone())
            torch.cuda.current_stream().wait_stream(side)
            self.train_graph = torch.cuda.CUDAGraph()
            self.optim.zero_grad(set_to_none=True)
            with torch.cuda.graph(self.train_graph):
                self.train_loss = self._train_step()
            self.eval_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.eval_graph), torch.no_grad():
                self.eval_out = self._forward()
            done = warm
        else:
            done = 0
        for _ in range(n_train - done):
            self.train_graph.replay()
            losses.append(self.train_loss.detach().reshape(()).clone())
        self.eval_graph.replay()
        loss, disp, poses, poses_inv = self.eval_out
        losses.append(loss.detach().reshape(()).clone())
        return {"losses": torch.stack(losses), "disparity": disp.detach().clone(),
                "poses": [p.detach().clone() for p in poses], "poses_inv": [p.detach().clone() for p in poses_inv]}
