"""torch.library registrations (custom_ops.py): schema / fake-tensor / autograd-registration checks with
torch.library.opcheck, tracing under FakeTensorMode, and agreement with the autograd.Function path -- on the
test-only emulator build (the operators themselves insist on CUDA tensors)."""
import pytest
import torch

import goldens
from emu_lib import emu
from goldens import Golden, rel_l2
from tcsfm_b200 import _cabi, custom_ops, losses, ops, stn  # noqa: F401


@pytest.fixture()
def emu_ops(monkeypatch):
    monkeypatch.setattr(ops, "lib", emu)
    monkeypatch.setattr(ops, "_require_cuda", lambda *a: None)
    monkeypatch.setattr(ops, "ARITH_FLAGS", _cabi.ARITH_CPU)


def leaf(t):
    return t.clone().detach().requires_grad_(True)


def _warp_args(g):
    fr = g.frames()
    kinv, proj = stn.projection_matrices(-fr["poses"][0], fr["K"])
    six = torch.cat([fr["target"], fr["sources"][0]], 1)
    return six[:, 3:6], fr["depths"][0], fr["depths"][1], kinv.contiguous(), proj.detach().contiguous()


def test_opcheck_inverse_warp2(emu_ops):
    g = Golden("small_b2_24x40")
    img, d0, d1, kinv, proj = _warp_args(g)
    args = (img, leaf(d0), leaf(d1), kinv, leaf(proj))
    torch.library.opcheck(torch.ops.tcsfm.inverse_warp2.default, args,
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))


def test_opcheck_ssim(emu_ops):
    x, y = torch.rand(2, 3, 12, 20), torch.rand(2, 3, 12, 20)
    torch.library.opcheck(torch.ops.tcsfm.ssim.default, (leaf(x), leaf(y)),
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))


def test_custom_op_matches_function_path(emu_ops):
    g = Golden("mid_b2_64x96")
    fr = g.frames()
    res = []
    for fn in (stn.inverse_warp2, stn.inverse_warp2_op):
        d0, d1, p0 = leaf(fr["depths"][0]), leaf(fr["depths"][1]), leaf(-fr["poses"][0])
        pim, vm, pd, cd = fn(fr["sources"][0], d0, d1, p0, fr["K"])
        ((pim * g.t("in/g_img")).sum() + (pd * g.t("in/g_pd")).sum() + (cd * g.t("in/g_cd")).sum()).backward()
        res.append((pim, vm, pd, cd, d0.grad, d1.grad, p0.grad))
    for a, b in zip(res[0][:4], res[1][:4]):
        assert torch.equal(a, b)
    for a, b in zip(res[0][4:], res[1][4:]):
        assert rel_l2(a, b) < 1e-5
    x, y = torch.rand(2, 3, 17, 33), torch.rand(2, 3, 17, 33)
    xa, xb = leaf(x), leaf(x)
    a, b = losses.SSIM_Loss()(xa, y), torch.ops.tcsfm.ssim(xb, y)
    assert torch.equal(a, b)
    a.sum().backward(); b.sum().backward()
    assert torch.equal(xa.grad, xb.grad)


def test_traces_with_fake_tensors():
    """Shape inference without running any kernel (what torch.compile / export do first)."""
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        img = torch.empty(4, 3, 192, 640)
        one = torch.empty(4, 1, 192, 640)
        out = torch.ops.tcsfm.inverse_warp2(img, one, one, torch.empty(4, 3, 3), torch.empty(4, 3, 4))
        assert [tuple(o.shape) for o in out] == [(4, 3, 192, 640)] + [(4, 1, 192, 640)] * 3
        assert tuple(torch.ops.tcsfm.ssim(img, img).shape) == (4, 3, 192, 640)
