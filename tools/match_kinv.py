"""Finds the ordering of the batched-LU arithmetic behind torch.linalg.inv_ex on 3x3 fp32 matrices (cuBLAS getrf +
getrs on the identity) that reproduces the probed inverses bit for bit (tools/probe_kinv.py).  fp32 operations are
emulated in float64 (products of two fp32 values are exact there; the second rounding of sums / quotients cannot
change an fp32 result except on measure-zero ties).  Usage: python tools/match_kinv.py gpurun_out/kinv_probe.npz"""
import itertools
import sys

import numpy as np

f32 = np.float32


def r(x):
    return x.astype(np.float32).astype(np.float64)


def mul(a, b):
    return r(a * b)


def sub(a, b):
    return r(a - b)


def fnma(a, b, c, fused):          # c - a*b
    return r(c - a * b) if fused else r(c - r(a * b))


def div(a, b, recip):
    if recip:
        return r(a * r(1.0 / b))
    return r(a / b)


def lu_inverse(A, scale_recip, upd_fused, l_fused, u_order, u_fused, u_recip):
    n = A.shape[0]
    a = A.astype(np.float64).copy()
    perm = np.tile(np.arange(3), (n, 1))
    rows = np.arange(n)
    for col in range(3):
        piv = col + np.argmax(np.abs(a[:, col:, col]), axis=1)
        # swap rows col <-> piv
        tmp = a[rows, col, :].copy(); a[rows, col, :] = a[rows, piv, :]; a[rows, piv, :] = tmp
        tp = perm[rows, col].copy(); perm[rows, col] = perm[rows, piv]; perm[rows, piv] = tp
        p = a[:, col, col]
        for i in range(col + 1, 3):
            a[:, i, col] = div(a[:, i, col], p, scale_recip)
        for i in range(col + 1, 3):
            for j in range(col + 1, 3):
                a[:, i, j] = fnma(a[:, i, col], a[:, col, j], a[:, i, j], upd_fused)
    # B = P * I
    B = np.zeros((n, 3, 3))
    for i in range(3):
        B[rows, i, perm[:, i]] = 1.0
    # forward substitution, unit lower
    y = B.copy()
    y[:, 1, :] = fnma(a[:, 1, 0:1], y[:, 0, :], y[:, 1, :], l_fused)
    y[:, 2, :] = fnma(a[:, 2, 0:1], y[:, 0, :], y[:, 2, :], l_fused)
    y[:, 2, :] = fnma(a[:, 2, 1:2], y[:, 1, :], y[:, 2, :], l_fused)
    x = np.zeros_like(y)
    x[:, 2, :] = div(y[:, 2, :], a[:, 2, 2:3], u_recip)
    t1 = fnma(a[:, 1, 2:3], x[:, 2, :], y[:, 1, :], u_fused)
    x[:, 1, :] = div(t1, a[:, 1, 1:2], u_recip)
    if u_order == "asc":
        t0 = fnma(a[:, 0, 1:2], x[:, 1, :], y[:, 0, :], u_fused)
        t0 = fnma(a[:, 0, 2:3], x[:, 2, :], t0, u_fused)
    else:
        t0 = fnma(a[:, 0, 2:3], x[:, 2, :], y[:, 0, :], u_fused)
        t0 = fnma(a[:, 0, 1:2], x[:, 1, :], t0, u_fused)
    x[:, 0, :] = div(t0, a[:, 0, 0:1], u_recip)
    return x.astype(np.float32)


def main():
    d = np.load(sys.argv[1])
    fams = sorted(k[:-3] for k in d.files if k.endswith("_in"))
    combos = list(itertools.product([False, True], [False, True], [False, True], ["asc", "desc"], [False, True], [False, True]))
    for fam in fams:
        A, want = d[fam + "_in"], d[fam + "_inv"]
        best = []
        for c in combos:
            got = lu_inverse(A, *c)
            same = (got.view(np.uint32) == want.view(np.uint32)) | ((got == 0) & (want == 0))
            exact_bits = (got.view(np.uint32) == want.view(np.uint32)).all(axis=(1, 2)).mean()
            best.append((same.all(axis=(1, 2)).mean(), exact_bits, c))
        best.sort(key=lambda t: -t[0])
        print(fam, "top:", [(round(a, 4), round(b, 4), c) for a, b, c in best[:4]])


main()
