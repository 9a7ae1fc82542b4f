"""Per-operator device timings (CUDA events, L2-defeating input rotation) with algorithmic GB/s:
inverse_warp2, SSIM_Loss, the PFT photometric block, the pair loss, get_smooth_loss.
Usage: python tools/bench_ops.py [B H W]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tcsfm_b200 import _raw, losses, ops, stn, synth  # noqa: E402
from tcsfm_b200._lib import lib  # noqa: E402

b, h, w = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (24, 192, 640)
dev = torch.device("cuda:0")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
n_sets = max(2, int(400e6 // (b * h * w * 4 * 10)) + 1)
sets = [synth.make_frames(b, h, w, seed=s, device=dev, intrinsics=synth.scaled_intrinsics(h, w)) for s in range(n_sets)]
npx = b * h * w


def timeit(fn, iters=40):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(10_000_000)        # ~5 ms: every launch below is queued before the first one starts (no host gaps in the interval)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


rows = []


def report(name, ms, bytes_per_px):
    gbs = npx * bytes_per_px / (ms * 1e-3) / 1e9
    rows.append({"op": name, "ms": round(ms, 4), "alg_bytes_per_px": bytes_per_px, "GBps": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak, 4)})


L = lib()
pre = []
for fr in sets:
    kinv, proj = stn.projection_matrices(-fr["poses"][0], fr["K"])
    six = torch.cat([fr["target"], fr["sources"][0]], 1)
    pre.append((six, kinv.contiguous(), proj.contiguous()))
g3 = torch.randn(b, 3, h, w, device=dev)
g1 = torch.randn(b, 1, h, w, device=dev)
g6 = torch.randn(b, 6, h, w, device=dev)


def f_warp(i):
    fr, (six, kinv, proj) = sets[i % n_sets], pre[i % n_sets]
    return _raw.warp_fwd(L, six[:, 3:6], fr["depths"][0], fr["depths"][1], kinv, proj, 0)


def f_warp_stack(i):
    fr, (six, kinv, proj) = sets[i % n_sets], pre[i % n_sets]
    return _raw.warp_fwd(L, six[:, 3:6], fr["depths"][0], fr["depths"][1], kinv, proj, 0, stack_target=six[:, 0:3])


def f_warp_bwd(i):
    fr, (six, kinv, proj) = sets[i % n_sets], pre[i % n_sets]
    return _raw.warp_bwd(L, six[:, 3:6], fr["depths"][0], fr["depths"][1], kinv, proj, g3, g1, g1, 0)


report("warp_fwd (inverse_warp2)", timeit(f_warp), 4 + 16 + 24)
report("warp_fwd + pose-net stack", timeit(f_warp_stack), 4 + 16 + 12 + 24 + 24)
report("warp_bwd", timeit(f_warp_bwd), 4 + 16 + 20 + 8)

outs = [f_warp(i) for i in range(n_sets)]


def f_ssim(i):
    fr = sets[i % n_sets]
    return _raw.ssim_fwd(L, fr["target"], outs[i % n_sets][0])


def f_ssim_bwd(i):
    fr = sets[i % n_sets]
    return _raw.ssim_bwd(L, fr["target"], outs[i % n_sets][0], g3, False, True)


report("ssim_fwd (3 ch)", timeit(f_ssim), 36)
report("ssim_bwd (3 ch, grad y)", timeit(f_ssim_bwd), 48)


def f_photo(i):
    fr, (six, kinv, proj) = sets[i % n_sets], pre[i % n_sets]
    o = outs[i % n_sets]
    return _raw.photo_fwd(L, six[:, 0:3], six[:, 3:6], o[0], o[2], o[3], 0.15, 0.85)


photo_out = [f_photo(i) for i in range(n_sets)]


def f_photo_bwd(i):
    six = pre[i % n_sets][0]
    o = outs[i % n_sets]
    return _raw.photo_bwd(L, six[:, 0:3], o[0], o[2], o[3], photo_out[i % n_sets][4], g1, g1, 0.15, 0.85)


report("photo_fwd (train_mono.py:84-92)", timeit(f_photo), 44 + 16)
report("photo_bwd", timeit(f_photo_bwd), 32 + 20)


def f_smooth(i):
    fr = sets[i % n_sets]
    return _raw.smooth_fwd(L, fr["disps"][0], fr["target"])


report("smooth_fwd (get_smooth_loss)", timeit(f_smooth), 4 + 4 + 12)


# PFT window reduction (optimizer.py:45-97): 2 sources, the five stacked [2*S*B',1,H,W] maps of one window minibatch
from tcsfm_b200 import _cabi  # noqa: E402
pb, ns = max(1, b // 4), 2
maps = [[torch.rand(2 * ns * pb, 1, h, w, device=dev) for _ in range(5)] for _ in range(n_sets)]
for m in maps:
    m[1].round_()
    m[3].round_()
pflags = _cabi.PFT_AUTOMASK | _cabi.PFT_INVERSE | _cabi.PFT_DEPTH_CONSIST
g_loss = torch.ones(1, device=dev)


def f_pft(i):
    return _raw.pft_reduce_fwd(L, *maps[i % n_sets], pb, ns, pflags, 0.14)


pft_out = [f_pft(i) for i in range(n_sets)]


def f_pft_bwd(i):
    _, sums, kept = pft_out[i % n_sets]
    return _raw.pft_reduce_bwd(L, kept, sums, g_loss, pb, ns, pflags, 0.14)


# bytes per pixel of one map element: fwd reads diff, valid, weight of both halves + auto_err (fwd) + auto_mask (inv);
# bwd reads the same and writes g_diff, g_weight
npx_full, npx = npx, 2 * ns * pb * h * w
report("pft_reduce_fwd (optimizer.py:45-86)", timeit(f_pft), 16)
report("pft_reduce_bwd", timeit(f_pft_bwd), 16 + 8)
npx = npx_full
print(json.dumps({"shape": [b, h, w], "peak_GBps": peak, "rows": rows}, indent=1))
