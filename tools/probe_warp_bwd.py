"""Times the standalone inverse_warp2 backward (tcsfm_warp_bwd) at B=24 192x640 with and without the
source-depth / source-image gradient scatters; used with tools/probe_variants_ops.sh to A/B tuning builds."""
import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from tcsfm_b200 import _raw, stn, synth
from tcsfm_b200._lib import lib
L = lib(); dev = torch.device('cuda:0'); b, h, w = 24, 192, 640
sets = [synth.make_frames(b, h, w, seed=s, device=dev, intrinsics=synth.scaled_intrinsics(h, w)) for s in range(6)]
pre = []
for fr in sets:
    kinv, proj = stn.projection_matrices(-fr["poses"][0], fr["K"])
    pre.append((torch.cat([fr["target"], fr["sources"][0]], 1), kinv, proj))
g3 = torch.randn(b, 3, h, w, device=dev); g1 = torch.randn(b, 1, h, w, device=dev)
def timeit(fn, iters=40):
    for i in range(5): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for name, kw in [("full", {}), ("no ref-depth grad", {"need_ref_depth_grad": False}), ("img grad too", {"need_img_grad": True})]:
    def f(i):
        fr, (six, kinv, proj) = sets[i % 6], pre[i % 6]
        return _raw.warp_bwd(L, six[:, 3:6], fr["depths"][0], fr["depths"][1], kinv, proj, g3, g1, g1, 0, **kw)
    print("%-20s %.4f ms" % (name, timeit(f)))
