"""GPU probe: rounding of the batch-1 3x3 @ 3x3 and 3x3 @ 3x4 products of the pose algebra."""
import json, torch
dev = torch.device("cuda:0")
torch.manual_seed(0)
def fma(a, b, c): return (a.double() * b.double() + c.double()).float()
def cands(A, X):
    n = X.shape[1]
    out = {}
    def build(f):
        return torch.stack([torch.stack([f(A[i, 0], A[i, 1], A[i, 2], X[0, j], X[1, j], X[2, j]) for j in range(n)]) for i in range(3)])
    out["fma012"] = build(lambda a0, a1, a2, x0, x1, x2: fma(a2, x2, fma(a1, x1, a0 * x0)))
    out["fma210"] = build(lambda a0, a1, a2, x0, x1, x2: fma(a0, x0, fma(a1, x1, a2 * x2)))
    out["nofma012"] = build(lambda a0, a1, a2, x0, x1, x2: (a0 * x0 + a1 * x1) + a2 * x2)
    out["nofma_0_12"] = build(lambda a0, a1, a2, x0, x1, x2: a0 * x0 + (a1 * x1 + a2 * x2))
    out["fma_then_add"] = build(lambda a0, a1, a2, x0, x1, x2: fma(a1, x1, a0 * x0) + a2 * x2)
    return out
res = {"33x33": {}, "33x34": {}, "33x33_strided": {}}
trials = 200
for name, n in (("33x33", 3), ("33x34", 4)):
    tot = {}
    for t in range(trials):
        A = torch.randn(1, 3, 3, device=dev); X = torch.randn(1, 3, n, device=dev)
        Y = (A @ X)[0]
        for k, v in cands(A[0], X[0]).items():
            tot[k] = tot.get(k, 0) + int((v != Y).sum())
    res[name] = tot
print(json.dumps(res, indent=1))
