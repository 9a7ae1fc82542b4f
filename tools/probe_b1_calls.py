"""GPU probe: rounding of the two k=3 bmm calls of inverse_warp2 exactly as the reference issues them
(K^-1 from .inverse(), rot = proj[:, :, :3] slice) for batch 1 at assorted sizes."""
import json, torch
dev = torch.device("cuda:0")
def fma(a, b, c): return (a.double() * b.double() + c.double()).float()
def cands(A, X):   # A [3,3], X [3,N] -> dict of [3,N]
    out = {}
    a = [[A[i, k].expand(X.shape[1]) for k in range(3)] for i in range(3)]
    out["fma012"] = torch.stack([fma(a[i][2], X[2], fma(a[i][1], X[1], a[i][0] * X[0])) for i in range(3)])
    out["fma210"] = torch.stack([fma(a[i][0], X[0], fma(a[i][1], X[1], a[i][2] * X[2])) for i in range(3)])
    out["nofma012"] = torch.stack([(a[i][0] * X[0] + a[i][1] * X[1]) + a[i][2] * X[2] for i in range(3)])
    out["nofma_0_12"] = torch.stack([a[i][0] * X[0] + (a[i][1] * X[1] + a[i][2] * X[2]) for i in range(3)])
    out["fma_mix_a"] = torch.stack([fma(a[i][1], X[1], a[i][0] * X[0]) + a[i][2] * X[2] for i in range(3)])
    out["fma_mix_b"] = torch.stack([a[i][0] * X[0] + fma(a[i][2], X[2], a[i][1] * X[1]) for i in range(3)])
    return out
res = {}
torch.manual_seed(0)
for (h, w) in ((50, 211), (60, 140), (189, 76), (150, 181), (179, 287), (192, 640), (64, 96), (24, 40), (256, 320)):
    HW = h * w
    K = torch.tensor([[370.7 * w / 640, 1.3, 313.1 * w / 640], [0, 367.1 * h / 192, 94.6 * h / 192], [0, 0, 1.0]], device=dev).unsqueeze(0)
    Kinv = K.inverse()
    rows = torch.arange(0, h).view(1, h, 1).expand(1, h, w).float().to(dev); cols = torch.arange(0, w).view(1, 1, w).expand(1, h, w).float().to(dev)
    grid = torch.stack((cols, rows, torch.ones(1, h, w, device=dev)), dim=1).expand(1, 3, h, w).reshape(1, 3, -1)
    ray = Kinv @ grid
    r = {"kinv_strides": list(Kinv.stride())}
    c = cands(Kinv[0], grid[0])
    r["ray"] = {k: int((v != ray[0]).sum()) for k, v in c.items()}
    depth = torch.rand(1, HW, device=dev) + 0.1
    cam = ray * depth.unsqueeze(1)
    pose = torch.randn(1, 3, 4, device=dev) * 0.1 + torch.eye(3, 4, device=dev)
    proj = K @ pose
    rot, tr = proj[:, :, :3], proj[:, :, -1:]
    pc = rot @ cam
    c = cands(rot[0], cam[0])
    r["pc_slice"] = {k: int((v != pc[0]).sum()) for k, v in c.items()}
    pc2 = rot.contiguous() @ cam
    r["pc_contig_equal_slice"] = bool(torch.equal(pc, pc2))
    res["%dx%d" % (h, w)] = r
print(json.dumps(res, indent=1))
