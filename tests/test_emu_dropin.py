"""Python-level wiring of the drop-ins (autograd functions, Compute_Loss
orchestration) exercised on the CPU by pointing the operators at the test-only
emulator build.  The package itself has no such path: ops.lib() loads only the
CUDA library and the operators reject CPU tensors (see test_no_cpu_fallback)."""
import pytest
import torch

import goldens
from emu_lib import emu
from goldens import Golden, rel_l2
from tcsfm_b200 import _cabi, losses, ops, stn


@pytest.fixture()
def emu_ops(monkeypatch):
    monkeypatch.setattr(ops, "lib", emu)
    monkeypatch.setattr(ops, "_require_cuda", lambda *a: None)
    monkeypatch.setattr(ops, "ARITH_FLAGS", _cabi.ARITH_CPU)


def leaf(t):
    return t.clone().detach().requires_grad_(True)


def test_no_cpu_fallback():
    x = torch.rand(1, 3, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        losses.SSIM_Loss()(x, x)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        stn.inverse_warp2(x, x[:, :1], x[:, :1], torch.zeros(1, 6), torch.eye(3).unsqueeze(0))


def test_check_sizes_assertions(emu_ops):
    x = torch.rand(1, 3, 8, 8)
    with pytest.raises(AssertionError, match="wrong size for depth"):
        stn.inverse_warp2(x, x, x[:, :1], torch.zeros(1, 6), torch.eye(3).unsqueeze(0))
    with pytest.raises(AssertionError, match="wrong size for img"):
        stn.inverse_warp2(x[:, :2], x[:, :1], x[:, :1], torch.zeros(1, 6), torch.eye(3).unsqueeze(0))
    with pytest.raises(NotImplementedError):
        stn.inverse_warp2(x, x[:, :1], x[:, :1], torch.zeros(1, 6), torch.eye(3).unsqueeze(0), 'border')


@pytest.mark.parametrize("case", goldens.CASES)
def test_inverse_warp2_autograd(emu_ops, case):
    g = Golden(case)
    fr = g.frames()
    d0, d1, p0 = leaf(fr["depths"][0]), leaf(fr["depths"][1]), leaf(-fr["poses"][0])
    pim, vm, pd, cd = stn.inverse_warp2(fr["sources"][0], d0, d1, p0, fr["K"], 'zeros')
    assert torch.equal(vm, g.t("warp/valid_mask")) and not vm.requires_grad
    ((pim * g.t("in/g_img")).sum() + (pd * g.t("in/g_pd")).sum() + (cd * g.t("in/g_cd")).sum()).backward()
    assert rel_l2(d0.grad, g.t("warp/g_depth")) < 2e-5
    assert rel_l2(d1.grad, g.t("warp/g_ref_depth")) < 2e-5
    assert rel_l2(p0.grad, g.t("warp/g_pose")) < 2e-4


def test_inverse_warp2_unused_outputs(emu_ops):
    g = Golden(goldens.CASES[0])
    fr = g.frames()
    d0 = leaf(fr["depths"][0])
    pim, _, _, _ = stn.inverse_warp2(fr["sources"][0], d0, fr["depths"][1], -fr["poses"][0], fr["K"])
    pim.sum().backward()           # projected/computed depth grads are None inside backward
    assert torch.isfinite(d0.grad).all()


@pytest.mark.parametrize("case", goldens.CASES)
def test_ssim_module(emu_ops, case):
    g = Golden(case)
    x, y = leaf(g.t("in/target")), leaf(g.t("in/source0"))
    s = losses.SSIM_Loss()(x, y)
    assert torch.equal(s, g.t("ssim/map"))
    (s * g.t("in/g_img")).sum().backward()
    assert rel_l2(x.grad, g.t("ssim/g_x")) < 2e-5 and rel_l2(y.grad, g.t("ssim/g_y")) < 2e-5


@pytest.mark.parametrize("case", goldens.CASES)
@pytest.mark.parametrize("tag", ["train", "full", "noauto"])
def test_compute_pairwise_loss(emu_ops, case, tag):
    g = Golden(case)
    fr = g.frames()
    mod = losses.Compute_Loss(goldens.PAIR_CFGS[tag])
    d0, d1, p0 = leaf(fr["depths"][0]), leaf(fr["depths"][1]), leaf(-fr["poses"][0])
    l_rep, l_dep, diff, vmask, none = mod.compute_pairwise_loss(fr["target"], fr["sources"][0], d0, d1, p0, fr["K"], 5)
    assert none is None and diff.shape == vmask.shape == (fr["target"].shape[0], 1) + fr["target"].shape[2:]
    assert torch.equal(vmask, g.t("pair_%s/valid_mask" % tag))
    assert (diff - g.t("pair_%s/diff_img" % tag)).abs().max() < 2e-6
    assert abs(float(l_rep) - float(g.t("pair_%s/l_reprojection" % tag))) <= 1e-5 * max(abs(float(l_rep)), 1e-12)
    obj = l_rep + (diff * g.t("in/g_diff")).sum()
    if torch.is_tensor(l_dep):
        assert abs(float(l_dep) - float(g.t("pair_%s/l_depth" % tag))) <= 1e-5 * max(abs(float(l_dep)), 1e-12)
        obj = obj + 0.5 * l_dep
    else:
        assert l_dep == 0
    obj.backward()
    assert rel_l2(d0.grad, g.t("pair_%s/g_depth" % tag)) < 1e-4
    assert rel_l2(p0.grad, g.t("pair_%s/g_pose" % tag)) < 1e-3
    if tag != "train":
        assert rel_l2(d1.grad, g.t("pair_%s/g_ref_depth" % tag)) < 1e-4


@pytest.mark.parametrize("case", goldens.CASES)
@pytest.mark.parametrize("tag", ["train", "full", "smooth"])
def test_compute_loss_forward_backward(emu_ops, case, tag):
    g = Golden(case)
    fr = g.frames()
    mod = losses.Compute_Loss(goldens.LOSS_CFGS[tag])
    disps = [leaf(d) for d in fr["disps"]]
    poses, poses_inv = [leaf(p) for p in fr["poses"]], [leaf(p) for p in fr["poses_inv"]]
    out = mod(fr["sources"], fr["target"], [poses, poses_inv], [[disps[0]], [disps[1]], [disps[2]]], fr["K"])
    for k in ("l_reconstruct_inverse", "l_reconstruct_forward", "l_depth", "l_smooth", "total"):
        assert out[k].shape == (1,), k
        a, b = float(out[k]), float(g.t("loss_%s/%s" % (tag, k)))
        assert abs(a - b) <= 1e-5 * max(abs(b), 1e-12), (k, a, b)
    out["total"].sum().backward()
    for j in range(3):
        ref_g = g.t("loss_%s/g_disp%d" % (tag, j))
        got = disps[j].grad if disps[j].grad is not None else torch.zeros_like(ref_g)
        assert rel_l2(got, ref_g) < 1e-4, j
    for j in range(2):
        for name, lst in (("g_pose%d", poses), ("g_pose_inv%d", poses_inv)):
            ref_g = g.t(("loss_%s/" % tag) + name % j)
            got = lst[j].grad if lst[j].grad is not None else torch.zeros_like(ref_g)
            assert rel_l2(got, ref_g) < 1e-3, (name, j)


def test_compute_loss_mixed_upstream(emu_ops):
    """Gradients through the individual terms and through `total` together (both upstream
    pointers of tcsfm_frame_bwd_prepare), and an in-place add on `total` like
    train_mono.py:181, against the oracle."""
    from oracle import ref_torch as O
    g = Golden(goldens.CASES[0])
    fr = g.frames()
    cfg = goldens.LOSS_CFGS["full"]
    grads = []
    for impl in ("ours", "oracle"):
        disps = [leaf(d) for d in fr["disps"]]
        poses, poses_inv = [leaf(p) for p in fr["poses"]], [leaf(p) for p in fr["poses_inv"]]
        args = (fr["sources"], fr["target"], [poses, poses_inv], [[disps[0]], [disps[1]], [disps[2]]], fr["K"])
        out = losses.Compute_Loss(cfg)(*args) if impl == "ours" else O.compute_loss(cfg, *args)
        extra = 0.01 * (poses[0] ** 2).sum().reshape(1)
        out["total"] += extra                                   # in place, as the trainer adds l_pose_consist
        (2.0 * out["total"] + 3.0 * out["l_depth"] - 0.5 * out["l_reconstruct_forward"]).sum().backward()
        grads.append([torch.zeros_like(t) if t.grad is None else t.grad.clone() for t in disps + poses + poses_inv])
    for a, b in zip(*grads):
        assert rel_l2(a, b) < 1e-3


def test_compute_loss_unused_terms_need_no_backward(emu_ops):
    g = Golden(goldens.CASES[0])
    fr = g.frames()
    disps = [leaf(d) for d in fr["disps"]]
    out = losses.Compute_Loss(goldens.LOSS_CFGS["full"])(
        fr["sources"], fr["target"], [fr["poses"], fr["poses_inv"]], [[disps[0]], [disps[1]], [disps[2]]], fr["K"])
    out["l_depth"].sum().backward()                             # only the [3] upstream, total's is absent
    assert all(d.grad is not None and torch.isfinite(d.grad).all() for d in disps)


@pytest.mark.parametrize("case", goldens.CASES)
def test_pft_path(emu_ops, case):
    """solve_pose_iteratively(return_errors=True) + compute_optimization_loss against
    the reference-generated fixture (train_mono.py:41-120, optimizer.py:45-97)."""
    from tcsfm_b200 import pft, synth, train_mono
    g = Golden(case)
    fr = g.frames()
    seed = {"small_b2_24x40": 1, "mid_b2_64x96": 2, "yaw_b2_32x48": 3}[case]
    net = synth.TinyPoseNet(seed=seed)
    dl = [leaf(d) for d in fr["depths"]]
    poses, poses_inv, outputs = train_mono.solve_pose_iteratively(3, dl, net, fr["target"], fr["sources"], fr["K"],
                                                                  return_errors=True)
    for side in ("fwd", "inv"):
        for k in ("valid_mask", "auto_mask"):
            assert torch.equal(outputs[side][k], g.t("pft/%s/%s" % (side, k))), (side, k)
        for k in ("diff_img", "weight_mask", "auto_mask_error"):
            assert (outputs[side][k] - g.t("pft/%s/%s" % (side, k))).abs().max() < 2e-6, (side, k)
        assert outputs[side]["poses"].shape == (2 * fr["target"].shape[0], 3, 6)
    for j in range(2):
        assert (poses[j] - g.t("pft/pose%d" % j)).abs().max() < 1e-7
        assert (poses_inv[j] - g.t("pft/pose_inv%d" % j)).abs().max() < 1e-7
    tdisp = leaf(fr["disps"][0])
    loss = pft.compute_optimization_loss(goldens.PFT_OPTIONS, fr["target"], tdisp, fr["disps"][0] * 0.9 + 0.02,
                                         outputs["fwd"], outputs["inv"])
    ref = float(g.t("pft/loss"))
    assert abs(float(loss.detach()) - ref) <= 1e-5 * abs(ref)
    loss.sum().backward()
    # The depth-init term is SSIM on two smooth, 0.9-correlated disparity maps: variances of
    # ~1e-5 come out of E[x^2] - mu^2 at ~0.3, so the reference's own fp32 gradient is ~6e-4
    # (rel-L2) away from its fp64 evaluation (see test_ssim_gradient_noise_floor).
    assert rel_l2(tdisp.grad, g.t("pft/g_tdisp")) < 5e-4
    for j in range(3):
        assert rel_l2(dl[j].grad, g.t("pft/g_depth%d" % j)) < 1e-4, j


@pytest.mark.parametrize("case", goldens.CASES)
def test_photometric_error(emu_ops, case):
    from tcsfm_b200 import pft
    g = Golden(case)
    fr = g.frames()
    with torch.no_grad():
        res = pft.compute_photometric_error(fr["target"][:1], fr["sources"][0][:1], fr["depths"][0][:1],
                                            fr["depths"][1][:1], fr["poses"][0][:1], fr["K"][:1])
    assert torch.equal(res["valid_mask"], g.t("photo/valid_mask"))
    for k in ("diff_img", "img_rec", "weight_mask"):
        assert (res[k] - g.t("photo/%s" % k)).abs().max() < 2e-6, k


def test_ssim_gradient_noise_floor(emu_ops):
    """On smooth, highly correlated inputs (the disparity-init term, optimizer.py:89-90) the
    kernel's gradient is closer to the reference's fp32 autograd than that is to fp64."""
    from oracle import ref_torch as O
    g = Golden("mid_b2_64x96")
    x = g.frames()["disps"][0]
    y = x * 0.9 + 0.02
    gout = torch.full_like(x, 0.1 / x.numel())

    def ref(dtype):
        xx = x.to(dtype).clone().requires_grad_(True)
        (O.ssim_dissimilarity(xx, y.to(dtype)) * gout.to(dtype)).sum().backward()
        return xx.grad
    r32, r64 = ref(torch.float32), ref(torch.float64)
    xx = leaf(x)
    (losses.SSIM_Loss()(xx, y) * gout).sum().backward()
    assert rel_l2(xx.grad, r32) < rel_l2(r32, r64)
    assert rel_l2(xx.grad, r64) < 1.5 * rel_l2(r32, r64)


def test_pft_driver_matches_oracle_backend(emu_ops):
    """Two optimisation epochs of a PFT window with stand-in networks: fused path (emulated) vs
    the oracle plugged into the same driver."""
    from oracle import ref_torch as O
    from tcsfm_b200 import pft_driver, synth

    class OracleBackend:
        solve_pose_iteratively = staticmethod(O.iterative_pose)
        compute_optimization_loss = staticmethod(O.pft_window_loss)
        disp_to_depth = staticmethod(O.disp_to_depth)

    fr = synth.make_frames(2, 24, 40, seed=6)
    depth_net, pose_net = synth.TinyDepthNet(seed=2), synth.TinyPoseNet(seed=2)
    opts = {"epochs": 3}
    got = pft_driver.optimize_window(depth_net, pose_net, fr["target"], fr["sources"], fr["K"], opts, iterations=2)
    ref = pft_driver.optimize_window(depth_net, pose_net, fr["target"], fr["sources"], fr["K"], opts, iterations=2,
                                     backend=OracleBackend)
    assert torch.allclose(got["losses"], ref["losses"], rtol=1e-5, atol=0), (got["losses"], ref["losses"])
    assert (got["disparity"] - ref["disparity"]).abs().max() < 1e-6
    # the caller's network is untouched: every window starts from the same weights
    assert all(torch.equal(a, b) for a, b in zip(depth_net.state_dict().values(), synth.TinyDepthNet(seed=2).state_dict().values()))


@pytest.mark.parametrize("shape", [(2, 24, 40), (3, 17, 33), (1, 2, 2)])
def test_smooth_loss_vs_oracle(emu_ops, shape):
    from oracle import ref_torch as O
    b, h, w = shape
    gen = torch.Generator().manual_seed(b * 100 + h)
    img = torch.rand(b, 3, h, w, generator=gen)
    d_ref = torch.rand(b, 1, h, w, generator=gen).add_(0.05).requires_grad_(True)
    d_got = d_ref.detach().clone().requires_grad_(True)
    ref = O.smooth_loss(d_ref, img)
    got = losses.get_smooth_loss(d_got, img)
    assert got.shape == ref.shape == ()
    assert abs(float(got) - float(ref)) <= 1e-5 * abs(float(ref))
    (ref * 1.7).backward()
    (got * 1.7).backward()
    assert rel_l2(d_got.grad, d_ref.grad) < 1e-4


def test_l1_only_configuration_matches_oracle(emu_ops):
    """l_ssim=False (no reference script uses it): diff_img keeps 3 channels; composed path."""
    from oracle import ref_torch as O
    g = Golden(goldens.CASES[0])
    fr = g.frames()
    cfg = dict(goldens.FULL_CFG, l_ssim=False)
    d0 = leaf(fr["depths"][0])
    got = losses.Compute_Loss(cfg).compute_pairwise_loss(fr["target"], fr["sources"][0], d0, fr["depths"][1],
                                                         -fr["poses"][0], fr["K"], 5)
    d0r = leaf(fr["depths"][0])
    ref = O.pairwise_loss(cfg, fr["target"], fr["sources"][0], d0r, fr["depths"][1], -fr["poses"][0], fr["K"])
    assert got[2].shape == ref[2].shape and got[2].shape[1] == 3
    assert torch.equal(got[3], ref[3]) and (got[2] - ref[2]).abs().max() < 1e-6
    (got[2].sum() + got[0]).backward()
    (ref[2].sum() + ref[0]).backward()
    assert rel_l2(d0.grad, d0r.grad) < 1e-4


@pytest.mark.parametrize("smooth", [False, True])
def test_compute_loss_multi_scale_vs_oracle(emu_ops, smooth):
    """num_scales = 3 with the lower-scale disparities at their own resolution (losses.py:86-87,
    102-103): the fused node upsamples inside its disp -> depth kernel."""
    from oracle import ref_torch as O
    from tcsfm_b200 import synth
    fr = synth.make_frames(2, 24, 40, seed=3)
    cfg = dict(goldens.LOSS_CFGS["full"], num_scales=3, l_smooth=smooth)
    gen = torch.Generator().manual_seed(5)
    sizes = [(24, 40), (12, 20), (5, 9)]
    base = [[torch.nn.functional.interpolate(d, s, mode="bilinear", align_corners=False) if s != (24, 40) else d
             for s in sizes] for d in fr["disps"]]
    results = []
    for impl in ("ours", "oracle"):
        disps = [[leaf(t) for t in per_frame] for per_frame in base]
        poses, poses_inv = [leaf(p) for p in fr["poses"]], [leaf(p) for p in fr["poses_inv"]]
        args = (fr["sources"], fr["target"], [poses, poses_inv], disps, fr["K"])
        out = losses.Compute_Loss(cfg)(*args) if impl == "ours" else O.compute_loss(cfg, *args)
        out["total"].sum().backward()
        grads = [t.grad if t.grad is not None else torch.zeros_like(t) for per_frame in disps for t in per_frame] + \
                [t.grad if t.grad is not None else torch.zeros_like(t) for t in poses + poses_inv]
        results.append(({k: float(v.detach()) for k, v in out.items()}, grads))
    (la, ga), (lb, gb) = results
    for k in lb:
        assert abs(la[k] - lb[k]) <= 1e-5 * max(abs(lb[k]), 1e-12), (k, la[k], lb[k])
    for a, b in zip(ga, gb):
        assert a.shape == b.shape and rel_l2(a, b) < 1e-3


@pytest.mark.parametrize("tag", ["train", "full"])
def test_compute_loss_single_source_vs_oracle(emu_ops, tag):
    """ScanNet-style pairs (one source image): the per-pixel min runs over a single candidate."""
    from oracle import ref_torch as O
    from tcsfm_b200 import synth
    fr = synth.make_frames(2, 24, 40, n_src=1, seed=8)
    cfg = goldens.LOSS_CFGS[tag]
    results = []
    for impl in ("ours", "oracle"):
        disps = [[leaf(d)] for d in fr["disps"]]
        poses, poses_inv = [leaf(p) for p in fr["poses"]], [leaf(p) for p in fr["poses_inv"]]
        args = (fr["sources"], fr["target"], [poses, poses_inv], disps, fr["K"])
        out = losses.Compute_Loss(cfg)(*args) if impl == "ours" else O.compute_loss(cfg, *args)
        out["total"].sum().backward()
        leaves = [d[0] for d in disps] + poses + poses_inv
        results.append(({k: float(v.detach()) for k, v in out.items()},
                        [t.grad if t.grad is not None else torch.zeros_like(t) for t in leaves]))
    (la, ga), (lb, gb) = results
    for k in lb:
        assert abs(la[k] - lb[k]) <= 1e-5 * max(abs(lb[k]), 1e-12), (k, la[k], lb[k])
    for a, b in zip(ga, gb):
        assert rel_l2(a, b) < 1e-3


@pytest.mark.parametrize("overrides", [{"l_inverse": False}, {"l_reconstruction": False, "l_smooth": True},
                                       {"l_inverse": False, "with_auto_mask": False, "l_depth_consist": False}])
def test_compute_loss_switches_vs_oracle(emu_ops, overrides):
    """Configuration switches the benchmark profiles leave on: forward direction only, smoothness only."""
    from oracle import ref_torch as O
    g = Golden(goldens.CASES[0])
    fr = g.frames()
    cfg = dict(goldens.LOSS_CFGS["full"], **overrides)
    results = []
    for impl in ("ours", "oracle"):
        disps = [[leaf(d)] for d in fr["disps"]]
        poses, poses_inv = [leaf(p) for p in fr["poses"]], [leaf(p) for p in fr["poses_inv"]]
        args = (fr["sources"], fr["target"], [poses, poses_inv], disps, fr["K"])
        out = losses.Compute_Loss(cfg)(*args) if impl == "ours" else O.compute_loss(cfg, *args)
        out["total"].sum().backward()
        leaves = [d[0] for d in disps] + poses + poses_inv
        results.append(({k: float(v.detach()) for k, v in out.items()},
                        [t.grad if t.grad is not None else torch.zeros_like(t) for t in leaves]))
    (la, ga), (lb, gb) = results
    for k in lb:
        assert abs(la[k] - lb[k]) <= 1e-5 * max(abs(lb[k]), 1e-12), (k, la[k], lb[k])
    for a, b in zip(ga, gb):
        assert rel_l2(a, b) < 1e-3 or (a.abs().max() == 0 and b.abs().max() == 0)


def test_compute_loss_three_sources_vs_oracle(emu_ops):
    """Three source frames: six pair groups in one launch and a three-way per-pixel min (the
    out-of-line routing path of the backward)."""
    from oracle import ref_torch as O
    from tcsfm_b200 import synth
    fr = synth.make_frames(2, 24, 40, n_src=3, seed=12)
    cfg = goldens.LOSS_CFGS["full"]
    results = []
    for impl in ("ours", "oracle"):
        disps = [[leaf(d)] for d in fr["disps"]]
        poses, poses_inv = [leaf(p) for p in fr["poses"]], [leaf(p) for p in fr["poses_inv"]]
        args = (fr["sources"], fr["target"], [poses, poses_inv], disps, fr["K"])
        out = losses.Compute_Loss(cfg)(*args) if impl == "ours" else O.compute_loss(cfg, *args)
        out["total"].sum().backward()
        leaves = [d[0] for d in disps] + poses + poses_inv
        results.append(({k: float(v.detach()) for k, v in out.items()},
                        [t.grad if t.grad is not None else torch.zeros_like(t) for t in leaves]))
    (la, ga), (lb, gb) = results
    for k in lb:
        assert abs(la[k] - lb[k]) <= 1e-5 * max(abs(lb[k]), 1e-12), (k, la[k], lb[k])
    for a, b in zip(ga, gb):
        assert rel_l2(a, b) < 1e-3


# ---- regressions for the round-1 review findings ------------------------------------------------

def test_two_scales_with_num_scales_one_matches_oracle(emu_ops):
    """A caller may hand more disparity scales than config['num_scales'] says: the reference iterates over
    every entry of `disparity` (losses.py:84) and divides by num_scales (losses.py:136), so `total` must be
    the sum over all passed scales (the fused per-scale total is only valid for a single scale)."""
    from oracle import ref_torch as O
    g = Golden(goldens.CASES[0])
    fr = g.frames()
    cfg = goldens.LOSS_CFGS["full"]
    low = [torch.nn.functional.avg_pool2d(d, 2) for d in fr["disps"]]
    outs = []
    for impl in ("ours", "oracle"):
        disps = [[leaf(d), leaf(l)] for d, l in zip(fr["disps"], low)]
        args = (fr["sources"], fr["target"], [fr["poses"], fr["poses_inv"]], disps, fr["K"])
        out = losses.Compute_Loss(cfg)(*args) if impl == "ours" else O.compute_loss(cfg, *args)
        out["total"].sum().backward()
        outs.append((out, [t.grad for ds in disps for t in ds]))
    for k in ("l_reconstruct_inverse", "l_reconstruct_forward", "l_depth", "total"):
        a, b = float(outs[0][0][k]), float(outs[1][0][k])
        assert abs(a - b) <= 1e-5 * abs(b), (k, a, b)
    for a, b in zip(outs[0][1], outs[1][1]):
        assert rel_l2(a, b) < 1e-3


def test_eight_wide_pose_vectors(emu_ops):
    """check_sizes accepts [B,8] poses (models/stn.py:252); the gradient comes back [B,8] with zeros in the
    two unused columns."""
    g = Golden(goldens.CASES[0])
    fr = g.frames()
    cfg = goldens.LOSS_CFGS["full"]

    def run(widen):
        pad = (lambda p: torch.cat([p, torch.ones(p.shape[0], 2)], 1)) if widen else (lambda p: p)
        poses, poses_inv = [leaf(pad(p)) for p in fr["poses"]], [leaf(pad(p)) for p in fr["poses_inv"]]
        disps = [leaf(d) for d in fr["disps"]]
        out = losses.Compute_Loss(cfg)(fr["sources"], fr["target"], [poses, poses_inv], [[d] for d in disps], fr["K"])
        out["total"].sum().backward()
        return float(out["total"]), poses[0].grad
    v6, g6 = run(False)
    v8, g8 = run(True)
    assert v6 == v8 and g8.shape == (g6.shape[0], 8)
    assert torch.equal(g8[:, 0:6], g6) and float(g8[:, 6:].abs().sum()) == 0


def test_shape_mismatches_are_rejected(emu_ops):
    """Depth / intrinsics tensors that do not match the images must raise instead of being read out of bounds."""
    g = Golden(goldens.CASES[0])
    fr = g.frames()
    mod = losses.Compute_Loss(goldens.LOSS_CFGS["full"])
    small = fr["depths"][0][:, :, ::2, ::2].contiguous()
    with pytest.raises(ValueError, match="tgt_depth"):
        mod.compute_pairwise_loss(fr["target"], fr["sources"][0], small, fr["depths"][1], -fr["poses"][0], fr["K"], 0)
    with pytest.raises(ValueError, match="ref_depth"):
        mod.compute_pairwise_loss(fr["target"], fr["sources"][0], fr["depths"][0], small, -fr["poses"][0], fr["K"], 0)
    with pytest.raises((ValueError, RuntimeError)):
        mod.compute_pairwise_loss(fr["target"], fr["sources"][0], fr["depths"][0], fr["depths"][1], -fr["poses"][0],
                                  fr["K"][:1], 0)


def test_image_gradients_are_refused(emu_ops):
    g = Golden(goldens.CASES[0])
    fr = g.frames()
    mod = losses.Compute_Loss(goldens.LOSS_CFGS["full"])
    with pytest.raises(NotImplementedError, match="images"):
        mod.compute_pairwise_loss(leaf(fr["target"]), fr["sources"][0], fr["depths"][0], fr["depths"][1],
                                  -fr["poses"][0], fr["K"], 0)
    with pytest.raises(NotImplementedError, match="images"):
        mod(fr["sources"], leaf(fr["target"]), [fr["poses"], fr["poses_inv"]], [[d] for d in fr["disps"]], fr["K"])


def test_intrinsics_inverse_is_never_stale():
    """K^-1 is recomputed per call (models/stn.py:257): views and in-place edits of K must be honoured."""
    k = torch.tensor([[[370.7, 0.3, 313.1], [0.1, 367.1, 94.6], [0.0, 0.0, 1.0]]]).repeat(2, 1, 1)
    a = stn.inverse_intrinsics(k)
    assert torch.equal(a, k.inverse())
    assert torch.equal(stn.inverse_intrinsics(k.transpose(1, 2)), k.transpose(1, 2).inverse())
    k.data[:, 0, 0] = 500.0
    assert torch.equal(stn.inverse_intrinsics(k), k.inverse())


PFT_VARIANTS = [
    {},                                                     # the scripts' defaults (run_sequential_optimization.py:69-99)
    {"diff_img_argmin": False},
    {"automasking": False},
    {"l_inverse_reconstruction": False},
    {"l_depth_consist": False},
    {"diff_img_argmin": False, "automasking": False, "l_depth_consist": False},
    {"l_depth_init": False, "l_smooth": True, "l_pose_consist": True},
]


@pytest.mark.parametrize("n_src", [1, 2, 3])
@pytest.mark.parametrize("variant", range(len(PFT_VARIANTS)))
def test_pft_reduce_option_matrix_vs_oracle(emu_ops, n_src, variant):
    """tcsfm_pft_reduce_fwd/bwd against the oracle's restatement of optimizer.py:45-97 for every option
    switch, 1-3 sources, maps handed over as slices of one stack (the zero-copy path) and as separate
    tensors (the concatenating path), with ties in the per-pixel minimum."""
    from oracle import ref_torch as O
    from tcsfm_b200 import pft
    opts = dict(goldens.PFT_OPTIONS, num_source_imgs=n_src, **PFT_VARIANTS[variant])
    gen = torch.Generator().manual_seed(100 * n_src + variant)
    bsz, h, w = 2, 9, 13
    rows = 2 * n_src * bsz
    diff = torch.rand(rows, 1, h, w, generator=gen)
    diff[bsz:2 * bsz] = torch.where(torch.rand(bsz, 1, h, w, generator=gen) < 0.3, diff[0:bsz], diff[bsz:2 * bsz])  # ties
    valid = (torch.rand(rows, 1, h, w, generator=gen) < 0.8).float()
    aerr = torch.rand(rows, 1, h, w, generator=gen)
    amask = (torch.rand(rows, 1, h, w, generator=gen) < 0.7).float()
    weight = torch.rand(rows, 1, h, w, generator=gen)
    target = torch.rand(bsz, 3, h, w, generator=gen)
    disp0 = torch.rand(bsz, 1, h, w, generator=gen)
    poses = torch.randn(rows, 3, 6, generator=gen) * 0.01
    split = n_src * bsz
    results = []
    for mode in ("stacked", "separate", "oracle"):
        d, wt, td = leaf(diff), leaf(weight), leaf(disp0)
        if mode == "separate":
            halves = lambda t: (t[:split].clone(), t[split:].clone())   # noqa: E731
        else:
            halves = lambda t: (t[:split], t[split:])                   # noqa: E731
        data = [{}, {}]
        for key, t in (("diff_img", d), ("valid_mask", valid), ("auto_mask_error", aerr), ("auto_mask", amask),
                       ("weight_mask", wt), ("poses", poses)):
            data[0][key], data[1][key] = halves(t)
        fn = O.pft_window_loss if mode == "oracle" else pft.compute_optimization_loss
        loss = fn(opts, target, td, disp0 * 0.9 + 0.02, data[0], data[1])
        loss.sum().backward()
        results.append((float(loss.detach()), d.grad, wt.grad, td.grad, loss.shape))
    ref = results[2]
    for got in results[:2]:
        assert got[4] == ref[4]                                  # [1] with the arg-min term, 0-d without
        assert abs(got[0] - ref[0]) <= 2e-6 * abs(ref[0]), (got[0], ref[0])
        assert rel_l2(got[1], ref[1]) < 1e-5
        if ref[2] is not None and float(ref[2].abs().sum()) > 0:
            assert rel_l2(got[2], ref[2]) < 1e-5
        if ref[3] is not None:
            assert rel_l2(got[3], ref[3]) < 5e-4


def test_ssim_mean_matches_map_mean(emu_ops):
    x, y = torch.rand(2, 1, 19, 70), torch.rand(2, 1, 19, 70)
    xa, xb = leaf(x), leaf(x)
    a = ops.SsimMeanFn.apply(xa, y)
    b = losses.SSIM_Loss()(xb, y).mean()
    assert abs(float(a) - float(b)) < 1e-6 * abs(float(b))
    (3.0 * a).sum().backward()
    (3.0 * b).backward()
    assert rel_l2(xa.grad, xb.grad) < 1e-5


def test_images_from_uint8_matches_the_loader_conversion(emu_ops):
    """utils/custom_transforms.py:74: torch.from_numpy(im).float() / 255 -- every byte value, odd lengths, views."""
    from tcsfm_b200 import dataformat
    u8 = torch.arange(0, 256, dtype=torch.uint8).repeat(5)[:1277].reshape(1, 1, 1, 1277)
    assert torch.equal(dataformat.images_from_uint8(u8), u8.float() / 255)
    img = torch.randint(0, 256, (2, 3, 13, 21), dtype=torch.uint8, generator=torch.Generator().manual_seed(0))
    out = torch.empty(2, 3, 13, 21)
    assert dataformat.images_from_uint8(img, out) is out and torch.equal(out, img.float() / 255)
    assert torch.equal(dataformat.images_from_uint8(img[:, :, 1:, 3:]), img[:, :, 1:, 3:].float() / 255)
    with pytest.raises(TypeError):
        dataformat.images_from_uint8(out)


def test_pack_slab_views_alias_one_buffer():
    """dataformat.pack_slab: the minibatch as one contiguous buffer (one host->device copy), views 256-byte aligned."""
    from tcsfm_b200 import dataformat, synth
    fr = synth.make_frames(2, 10, 14, seed=1)
    tensors = {"target": fr["target"], "K": fr["K"], "pose0": fr["poses"][0], "disp0": fr["disps"][0]}
    slab, views = dataformat.pack_slab(tensors)
    assert slab.dim() == 1 and slab.dtype == torch.float32
    for name, t in tensors.items():
        v = views[name]
        assert v.shape == t.shape and torch.equal(v, t)
        off = (v.data_ptr() - slab.data_ptr()) // 4
        assert off % 64 == 0 and 0 <= off and off + t.numel() <= slab.numel()
    slab.zero_()
    assert all(float(v.abs().sum()) == 0.0 for v in views.values())          # views alias the slab
    other, views2 = dataformat.pack_slab(tensors)
    slab.copy_(other)                                                         # "one copy" moves every tensor
    assert all(torch.equal(views[k], tensors[k]) for k in tensors)
    with pytest.raises(TypeError):
        dataformat.pack_slab({"a": torch.zeros(3), "b": torch.zeros(3, dtype=torch.int32)})


def test_compute_loss_rejects_a_pose_batch_that_disagrees_with_the_intrinsics(emu_ops):
    """The fused frame node reads the poses in place through a pointer table: a pose tensor with fewer rows than the
    intrinsics must be refused before any kernel sees it (the reference fails with a shape error as well)."""
    from tcsfm_b200 import synth
    fr = synth.make_frames(2, 24, 40, seed=3)
    loss = losses.Compute_Loss(goldens.FULL_CFG)
    short = [p[:1] for p in fr["poses"]]
    with pytest.raises((ValueError, RuntimeError, AssertionError)):
        loss(fr["sources"], fr["target"], [short, fr["poses_inv"]], [[d] for d in fr["disps"]], fr["K"])
