"""GPU probe: bisect the column count at which torch's [1,3,3]@[1,3,HW] switches kernels."""
import json, torch
dev = torch.device("cuda:0")
torch.manual_seed(0)
def fma(a, b, c): return (a.double() * b.double() + c.double()).float()
def kind(HW):
    A = torch.randn(1, 3, 3, device=dev); X = torch.randn(1, 3, HW, device=dev)
    Y = (A @ X)[0]; x = X[0]
    bf = bn = 0
    for i in range(3):
        a = [A[0, i, k].expand(HW) for k in range(3)]
        f = fma(a[2], x[2], fma(a[1], x[1], a[0] * x[0]))
        nf = (a[0] * x[0] + a[1] * x[1]) + a[2] * x[2]
        bf += int((f != Y[i]).sum()); bn += int((nf != Y[i]).sum())
    return "fma" if bf == 0 else ("nofma" if bn == 0 else "other")
lo, hi = 230000, 262143
assert kind(lo) == "nofma" and kind(hi) != "nofma", (kind(lo), kind(hi))
while hi - lo > 1:
    mid = (lo + hi) // 2
    if kind(mid) == "nofma": lo = mid
    else: hi = mid
out = {"last_nofma": lo, "first_other": hi, "kind_first_other": kind(hi)}
# is it monotone / alignment dependent?  sample around and far
out["samples"] = {str(n): kind(n) for n in (lo - 7, lo - 1, lo, hi, hi + 1, hi + 7, 245760, 250000, 255000, 260000, 262143, 262144, 262145, 300001)}
# does the leading dims (3x3 @ 3xHW inside inverse_warp2: expand+reshape copy) matter? same op via matmul on non-contiguous input
print(json.dumps(out, indent=1))
