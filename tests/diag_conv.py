"""Diagnostic (not a pytest file): structured single-tile runs of gd_conv_igemm that reveal layout / descriptor
mistakes.  Usage on the GPU box: python tests/diag_conv.py"""
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th

from guided_diffusion_clip_b200 import _lib as L
from guided_diffusion_clip_b200.engine import pack_1x1, pack_conv3x3
from tests import gpu_helpers as H


def where_from(out, a, m, j):
    """Find which A[m', k'] equals out[m, j] (A has unique values)."""
    hit = (a == out[m, j]).nonzero()
    return hit[:3].tolist()


def gemm_identity(cout=64, cin=64, h=8, w=16, n=1):
    print(f"--- 1x1 GEMM with W = I: n={n} h={h} w={w} cin={cin} cout={cout}")
    m = n * h * w
    a = (th.arange(m * cin, dtype=th.float32).reshape(m, cin) % 2039) / 8.0  # unique-ish, exactly representable
    a = a.cuda()
    x = a.reshape(n, h, w, cin).half().contiguous()
    wt = th.zeros(cout, cin)
    for i in range(min(cout, cin)):
        wt[i, i] = 1.0
    out = H.conv_igemm(x, cin, 0, pack_1x1(wt.cuda().reshape(cout, cin, 1)), None, cout, n, h, w, taps=1)
    th.cuda.synchronize()
    o = out.reshape(m, cout).float()
    exp = x.reshape(m, cin).float()[:, :cout]
    bad = (o != exp)
    print("mismatches:", int(bad.sum()), "of", o.numel())
    if bad.any():
        rows = bad.any(1).nonzero().flatten()[:6].tolist()
        for r in rows:
            cols = bad[r].nonzero().flatten()[:6].tolist()
            print(f" row {r}: bad cols {cols}; got {o[r, cols[:4]].tolist()} exp {exp[r, cols[:4]].tolist()}",
                  "from", [where_from(o, x.reshape(m, cin).float(), r, c) for c in cols[:2]])
        print(" row-wise bad counts (first 16 rows):", bad.sum(1)[:16].tolist())
        print(" col-wise bad counts (first 16 cols):", bad.sum(0)[:16].tolist())
    return int(bad.sum())


def conv_delta(h=8, w=16, cin=64, cout=64):
    print(f"--- 3x3 conv, one-hot weight per tap: h={h} w={w}")
    n = 1
    x = th.zeros(n, h, w, cin)
    for y in range(h):
        for xx in range(w):
            x[0, y, xx, :] = y * 16 + xx + 1
    x = x.half().cuda()
    total_bad = 0
    for tap in range(9):
        wt = th.zeros(cout, cin, 3, 3)
        wt[0, 0, tap // 3, tap % 3] = 1.0
        out = H.conv_igemm(x, cin, 0, pack_conv3x3(wt.cuda()), None, cout, n, h, w)
        th.cuda.synchronize()
        got = out[0, :, :, 0].float().cpu()
        dy, dx = tap // 3 - 1, tap % 3 - 1
        exp = th.zeros(h, w)
        for y in range(h):
            for xx in range(w):
                yy, xc = y + dy, xx + dx
                if 0 <= yy < h and 0 <= xc < w:
                    exp[y, xx] = yy * 16 + xc + 1
        bad = int((got != exp).sum())
        total_bad += bad
        print(f" tap {tap} (dy={dy},dx={dx}): mismatches {bad}")
        if bad:
            print("  got row0:", got[0].tolist())
            print("  exp row0:", exp[0].tolist())
            print("  got row1:", got[1].tolist())
    return total_bad


if __name__ == "__main__":
    print(th.cuda.get_device_name(0))
    bad = gemm_identity()
    bad += gemm_identity(cout=256, cin=128, h=16, w=16)
    bad += conv_delta()
    print("TOTAL BAD", bad)
