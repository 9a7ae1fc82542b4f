#!/bin/bash
# A/B timing of tuning builds of the stand-alone warp backward (same sources, -D macros) on the GPU.
# Usage (on the GPU box): bash tools/bench_warp_variants.sh "NAME:-DX=1,-DY=2" ...
mkdir -p build gpurun_out
for spec in "$@"; do
  name=${spec%%:*}; defs=${spec#*:}
  python - "$name" "$defs" <<'PY'
import sys
sys.path.insert(0, ".")
from tcsfm_b200 import build
name, defs = sys.argv[1], sys.argv[2]
build.build_variant("build/libtcsfm_%s.so" % name, [d[2:] for d in defs.split(",") if d.startswith("-D")])
PY
done
for lib in default $(ls build/libtcsfm_*.so 2>/dev/null); do
  if [ "$lib" != default ]; then export TCSFM_B200_LIB=$PWD/$lib; else unset TCSFM_B200_LIB; fi
  python - "$lib" <<'PY'
import sys
import torch
sys.path.insert(0, ".")
from tcsfm_b200 import _raw, stn, synth
from tcsfm_b200._lib import lib
b, h, w = 24, 192, 640
dev = torch.device("cuda:0")
sets = [synth.make_frames(b, h, w, seed=s, device=dev, intrinsics=synth.scaled_intrinsics(h, w)) for s in range(4)]
pre = []
for fr in sets:
    kinv, proj = stn.projection_matrices(-fr["poses"][0], fr["K"])
    pre.append((torch.cat([fr["target"], fr["sources"][0]], 1), kinv.contiguous(), proj.contiguous()))
g3, g1, g6 = (torch.randn(b, c, h, w, device=dev) for c in (3, 1, 6))
L = lib()
def timeit(fn, iters=40):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(10_000_000)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
def full(i):
    fr, (six, kinv, proj) = sets[i % 4], pre[i % 4]
    return _raw.warp_bwd(L, six[:, 3:6], fr["depths"][0], fr["depths"][1], kinv, proj, g3, g1, g1, 0)
def pft(i):      # non-final egomotion iteration: upstream of the image and of the pose-net stack only
    fr, (six, kinv, proj) = sets[i % 4], pre[i % 4]
    return _raw.warp_bwd(L, six[:, 3:6], fr["depths"][0], fr["depths"][1], kinv, proj, g3, None, None, 0, need_ref_depth_grad=False, g_stack=g6)
t_full, t_pft = timeit(full), timeit(pft)
npx = b * h * w
print("%-36s warp_bwd all grads %.4f ms (%.3f of peak)   image+stack grads %.4f ms" % (sys.argv[1], t_full, npx * 48 / t_full / 1e6 / 6529.7, t_pft))
PY
done
