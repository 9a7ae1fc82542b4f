"""GPU probe: accumulation order of eager torch.bmm / matmul for k=3 as a function of
the batch size (cuBLAS picks different kernels for batch 1)."""
import itertools, json, torch
dev = torch.device("cuda:0")
torch.manual_seed(0)
def fma(a, b, c): return (a.double() * b.double() + c.double()).float()
def mism(a, b): return float((a != b).float().mean())
out = {}
for B in (1, 2, 3, 4, 8, 24, 32):
    for HW in (960, 81920, 122880, 466992):
        A = torch.randn(B, 3, 3, device=dev); X = torch.randn(B, 3, HW, device=dev)
        Y = A @ X
        a = [[A[:, i, k].unsqueeze(1).expand(B, HW).contiguous() for k in range(3)] for i in range(3)]
        res = {}
        for order in itertools.permutations(range(3)):
            bad = 0.0
            for i in range(3):
                acc = None
                for k in order:
                    acc = a[i][k] * X[:, k] if acc is None else fma(a[i][k], X[:, k], acc)
                bad += mism(acc, Y[:, i]) / 3
            res["fma%d%d%d" % order] = round(bad, 4)
        bad = 0.0
        for i in range(3):
            acc = (a[i][0] * X[:, 0] + a[i][1] * X[:, 1]) + a[i][2] * X[:, 2]
            bad += mism(acc, Y[:, i]) / 3
        res["nofma_asc"] = round(bad, 4)
        # exact fp64 then round (e.g. tensor-core / wider accumulate)
        exact = (A.double() @ X.double()).float()
        res["fp64_round"] = round(mism(exact, Y), 4)
        res["bmm_vs_matmul_expand"] = round(mism(torch.bmm(A, X), Y), 4)
        out["B%d_HW%d" % (B, HW)] = {k: v for k, v in res.items() if v < 0.05} or res
# 3x3 @ 3x4 products of the pose algebra
for B in (1, 2, 8, 32):
    K = torch.randn(B, 3, 3, device=dev); T = torch.randn(B, 3, 4, device=dev)
    Y = K @ T
    res = {}
    for order in itertools.permutations(range(3)):
        acc = None
        for k in order:
            t = K[:, :, k:k + 1] * T[:, k:k + 1, :]
            acc = t if acc is None else fma(K[:, :, k:k + 1].expand(B, 3, 4), T[:, k:k + 1, :].expand(B, 3, 4), acc)
        res["fma%d%d%d" % order] = round(mism(acc, Y), 4)
    out["K@T_B%d" % B] = res
print(json.dumps(out, indent=1))
