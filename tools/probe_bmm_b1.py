"""GPU probe: for batch 1, at which HW does torch's [1,3,3]@[1,3,HW] switch from the
non-fused (a0*b0 + a1*b1) + a2*b2 kernel to the FMA chain?"""
import json, torch
dev = torch.device("cuda:0")
torch.manual_seed(0)
def fma(a, b, c): return (a.double() * b.double() + c.double()).float()
out = {}
sizes = [64, 960, 4096, 16384, 65536, 81920, 114688, 122880, 131072, 150000, 180000, 200000, 230000, 262144, 300000,
         350000, 400000, 428032, 466992, 479232, 600000, 1000000]
for HW in sizes:
    A = torch.randn(1, 3, 3, device=dev); X = torch.randn(1, 3, HW, device=dev)
    Y = (A @ X)[0]
    a = [[A[0, i, k].expand(HW).contiguous() for k in range(3)] for i in range(3)]
    x = X[0]
    bad_fma = bad_nofma = 0.0
    for i in range(3):
        f = fma(a[i][2], x[2], fma(a[i][1], x[1], a[i][0] * x[0]))
        nf = (a[i][0] * x[0] + a[i][1] * x[1]) + a[i][2] * x[2]
        bad_fma += float((f != Y[i]).float().mean()) / 3
        bad_nofma += float((nf != Y[i]).float().mean()) / 3
    out[HW] = "fma" if bad_fma == 0 else ("nofma" if bad_nofma == 0 else "other(%.3f,%.3f)" % (bad_fma, bad_nofma))
# batch 2 sanity at the same sizes
for HW in (64, 960, 122880):
    A = torch.randn(2, 3, 3, device=dev); X = torch.randn(2, 3, HW, device=dev)
    Y = A @ X
    f = fma(A[:, :, 2:3], X[:, 2:3], fma(A[:, :, 1:2], X[:, 1:2], A[:, :, 0:1] * X[:, 0:1]))
    out["B2_%d" % HW] = "fma" if bool((f == Y).all()) else "other"
print(json.dumps(out, indent=1))
