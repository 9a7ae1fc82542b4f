"""GPU probe for the bit-parity hazards of SURVEY.md App. B: which rounding
sequence do eager PyTorch's CUDA operators perform for the k=3 bmm, the
scalar true-divide, mean(dim=1), avg_pool2d and grid_sample?  Prints one JSON
object; run on the B200 box (`gpurun -- python tools/probe_arith.py`).  The result
is committed under profiles/ and is what csrc/tcsfm_math.cuh is written against."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

dev = torch.device("cuda:0")
torch.manual_seed(0)
out = {"torch": torch.__version__, "gpu": torch.cuda.get_device_name(0),
       "allow_tf32_matmul": torch.backends.cuda.matmul.allow_tf32}


def fma(a, b, c):     # fp32 fma emulated in fp64 (double rounding is ~1e-9 rare)
    return (a.double() * b.double() + c.double()).float()


def mism(a, b):
    return float((a != b).float().mean())


# ---- 1. bmm with k=3 (models/stn.py:47,210) -----------------------------------
for (B, HW) in ((8, 192 * 640), (32, 192 * 640), (16, 256 * 320), (2, 24 * 40)):
    A = torch.randn(B, 3, 3, device=dev)
    X = torch.randn(B, 3, HW, device=dev)
    Y = A @ X
    a = [[A[:, i, k].unsqueeze(1).expand(B, HW).contiguous() for k in range(3)] for i in range(3)]
    res = {}
    for name, order, first_mul in (("fma_asc", (0, 1, 2), True), ("fma_desc", (2, 1, 0), True)):
        bad = 0.0
        for i in range(3):
            acc = None
            for k in order:
                acc = a[i][k] * X[:, k] if acc is None else fma(a[i][k], X[:, k], acc)
            bad += mism(acc, Y[:, i]) / 3
        res[name] = bad
    bad = 0.0
    for i in range(3):
        acc = (a[i][0] * X[:, 0] + a[i][1] * X[:, 1]) + a[i][2] * X[:, 2]
        bad += mism(acc, Y[:, i]) / 3
    res["nofma_asc"] = bad
    out["bmm_k3_B%d_HW%d" % (B, HW)] = res

# K^-1-shaped A (exact zeros), pixel grid X as in pixel2cam
B, H, W = 8, 192, 640
K = torch.tensor([[370.7, 0, 313.1], [0, 367.1, 94.6], [0, 0, 1.0]], device=dev).repeat(B, 1, 1)
Kinv = K.inverse()
out["Kinv_row0"] = [float(v) for v in Kinv[0].flatten()]
out["Kinv_candidates"] = {"1/fx": float(1 / K[0, 0, 0]), "-cx/fx": float(-K[0, 0, 2] / K[0, 0, 0]),
                          "-(cx*(1/fx))": float(-(K[0, 0, 2] * (1 / K[0, 0, 0])))}
out["Kinv_cpu_equal"] = bool(torch.equal(Kinv.cpu(), K.cpu().inverse()))

# ---- 2. tensor / python scalar (models/stn.py:221) ---------------------------
x = torch.randn(1 << 20, device=dev)
d = 639
out["div_scalar"] = {"true_div": mism(x / d, x / torch.tensor(float(d), device=dev)),
                     "mul_recip_fp32": mism(x / d, x * (torch.tensor(1.0, device=dev) / torch.tensor(float(d), device=dev))),
                     "mul_recip_fp64": mism(x / d, x * torch.tensor(1.0 / d, device=dev, dtype=torch.float32))}

# ---- 3. mean(dim=1) over 3 channels (losses.py:158) ----------------------------
t = torch.rand(8, 3, 192, 640, device=dev)
m = t.mean(dim=1)
s = (t[:, 0] + t[:, 1]) + t[:, 2]
third = torch.tensor(1.0, device=dev) / torch.tensor(3.0, device=dev)
out["mean3"] = {"sum_div3": mism(s / torch.tensor(3.0, device=dev), m), "sum_mul_third": mism(s * third, m),
                "alt_order_mul_third": mism(((t[:, 1] + t[:, 2]) + t[:, 0]) * third, m),
                "keepdim_same": bool(torch.equal(t.mean(dim=1, keepdim=True)[:, 0], m))}

# ---- 4. avg_pool2d(3,1) (losses.py:16-20) --------------------------------------
p = torch.rand(2, 3, 66, 130, device=dev)
ap = F.avg_pool2d(p, 3, 1)
acc = torch.zeros_like(ap)
for dy in range(3):
    for dx in range(3):
        acc = acc + p[:, :, dy:dy + 64, dx:dx + 128]
out["avg_pool"] = {"rowmajor_div9": mism(acc / torch.tensor(9.0, device=dev), ap),
                   "rowmajor_mul_ninth": mism(acc * (torch.tensor(1.0, device=dev) / 9), ap)}

# ---- 5. the kernels against the eager CUDA oracle -------------------------------
from oracle import ref_torch as O                     # noqa: E402
from tcsfm_b200 import losses, stn, synth, ops, _cabi  # noqa: E402

for (b, h, w, yaw) in ((4, 192, 640, 0.01), (4, 256, 320, 0.05), (2, 376, 1242, 0.02)):
    fr = synth.make_frames(b, h, w, seed=7, yaw=yaw, device=dev,
                           intrinsics=synth.scaled_intrinsics(h, w))
    args = (fr["sources"][0], fr["depths"][0], fr["depths"][1], -fr["poses"][0], fr["K"])
    ref = O.inverse_warp2(*args)
    got = stn.inverse_warp2(*args)
    r = {}
    for name, a, bb in zip(("projected_img", "valid_mask", "projected_depth", "computed_depth"), got, ref):
        r[name] = {"mismatch_frac": mism(a, bb), "max_abs": float((a - bb).abs().max())}
    # intermediate coordinates of the oracle, to localise a mismatch
    cam = O.backproject(fr["depths"][0].squeeze(1), fr["K"].inverse())
    r["valid_frac"] = float(ref[1].mean())
    s_ref = O.ssim_dissimilarity(fr["target"], fr["sources"][0])
    s_got = losses.SSIM_Loss()(fr["target"], fr["sources"][0])
    r["ssim"] = {"mismatch_frac": mism(s_got, s_ref), "max_abs": float((s_got - s_ref).abs().max())}
    cfg = dict(O.DEFAULT_LOSS_CONFIG, l_depth_consist=True, with_depth_mask=True)
    pr = O.pairwise_loss(cfg, fr["target"], fr["sources"][0], fr["depths"][0], fr["depths"][1], -fr["poses"][0], fr["K"])
    pg = losses.Compute_Loss(cfg).compute_pairwise_loss(fr["target"], fr["sources"][0], fr["depths"][0],
                                                        fr["depths"][1], -fr["poses"][0], fr["K"], 5)
    r["pair_full"] = {"mask_mismatch_px": int((pg[3] != pr[3]).sum()), "mask_frac": float(pr[3].mean()),
                      "diff_mismatch_frac": mism(pg[2], pr[2]), "diff_max_abs": float((pg[2] - pr[2]).abs().max()),
                      "l_rep": [float(pg[0]), float(pr[0])], "l_dep": [float(pg[1]), float(pr[1])]}
    # CPU flavour against the CPU oracle
    ops.ARITH_FLAGS = _cabi.ARITH_CPU
    got_c = stn.inverse_warp2(*args)
    ref_c = O.inverse_warp2(*[a.cpu() for a in args])
    r["cpu_flavour_vs_cpu_oracle"] = {
        name: {"mismatch_frac": mism(a.cpu(), bb), "max_abs": float((a.cpu() - bb).abs().max())}
        for name, a, bb in zip(("projected_img", "valid_mask", "projected_depth", "computed_depth"), got_c, ref_c)}
    ops.ARITH_FLAGS = 0
    r["cuda_oracle_vs_cpu_oracle_valid_mismatch_px"] = int((ref[1].cpu() != ref_c[1]).sum())
    out["kernels_%dx%dx%d" % (b, h, w)] = r

print(json.dumps(out, indent=1))
