"""Property-style sweep of the pair kernels on the CPU emulator: random image sizes (tile edges,
widths that are and are not multiples of four, i.e. both staging paths of the backward), random
flag combinations and poses (sizes above the handful of pixels where the CPU BLAS behind the oracle changes its k=3 rounding); forward bit for bit, gradients within tolerance, against the oracle."""
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

import goldens
from emu_lib import emu
from goldens import rel_l2
from oracle import ref_torch as O
from tcsfm_b200 import _cabi, _raw, stn, synth

CPU = _cabi.ARITH_CPU


@settings(max_examples=30, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(h=st.integers(6, 40), w=st.integers(8, 140), seed=st.integers(0, 1000),
       auto_mask=st.booleans(), depth_mask=st.booleans(), depth_consist=st.booleans())
def test_pair_loss_random_shapes_and_flags(h, w, seed, auto_mask, depth_mask, depth_consist):
    fr = synth.make_frames(2, h, w, seed=seed, yaw=0.01 * (seed % 4), intrinsics=synth.scaled_intrinsics(max(h, 8), max(w, 8)))
    cfg = dict(goldens.FULL_CFG, with_auto_mask=auto_mask, with_depth_mask=depth_mask, l_depth_consist=depth_consist)
    flags = CPU | _cabi.SSIM
    flags |= _cabi.AUTO_MASK if auto_mask else 0
    flags |= _cabi.DEPTH_MASK if depth_mask else 0
    flags |= _cabi.DEPTH_CONSIST if depth_consist else 0
    pose = -fr["poses"][0]
    kinv, proj = stn.projection_matrices(pose, fr["K"])
    batch = _raw.PairBatch([{"tgt_img": fr["target"], "ref_img": fr["sources"][0], "tgt_depth": fr["depths"][0],
                             "ref_depth": fr["depths"][1], "kinv": kinv, "proj": proj}])
    diff, mask, sums, coef = _raw.pair_loss_fwd(emu(), batch, 0.15, 0.85, flags)
    td, rd = fr["depths"][0].clone().requires_grad_(True), fr["depths"][1].clone().requires_grad_(True)
    _, _, rdiff, rmask, _ = O.pairwise_loss(cfg, fr["target"], fr["sources"][0], td, rd, pose, fr["K"])
    assert torch.equal(mask[0], rmask)
    assert torch.equal(diff[0], rdiff)
    gen = torch.Generator().manual_seed(seed)
    g_diff = torch.randn(1, 2, 1, h, w, generator=gen)
    g_diff[0, :, :, :, : w // 3] = 0                      # dead upstream: exercises the zero-fill staging
    g_td, g_rd, _ = _raw.pair_loss_bwd(emu(), batch, mask, sums, coef, g_diff, None, 0.15, 0.85, flags, True)
    (rdiff * g_diff[0]).sum().backward()
    assert rel_l2(g_td[0], td.grad) < 1e-4
    if rd.grad is not None:
        assert rel_l2(g_rd[0], rd.grad) < 1e-4
