"""N-tile width for the small-spatial layers (8x8, 16x16 with 1024 channels): few pixel tiles, long K loops.
  python profiles/conv_small_probe.py            -> TFLOP/s per (batch, resolution, C_in, N tile)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th  # noqa: E402

from guided_diffusion_clip_b200 import _lib as L  # noqa: E402
from guided_diffusion_clip_b200.engine import pack_conv3x3  # noqa: E402
from tests import gpu_helpers as H  # noqa: E402

lib = L.load()


def run(n, hw, cin, cout, bn, reps=20):
    x = th.randn((n, hw, hw, cin), device="cuda").half()
    pack = pack_conv3x3(th.randn((cout, cin, 3, 3), device="cuda") * 0.01)
    b = th.zeros(cout, device="cuda")
    out = th.empty((n, hw, hw, cout), dtype=th.float16, device="cuda")
    f = lambda: H.conv_igemm(x, cin, 0, pack, b, cout, n, hw, hw, out_buf=out, bn=bn)  # noqa: E731
    f()
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    th.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms, 2.0 * n * hw * hw * cout * 9 * cin / ms / 1e9


for n in (8, 64):
    for hw, cin in ((8, 1024), (8, 2048), (16, 1024), (16, 2048), (32, 512)):
        row = []
        for bn in (256, 128, 64):
            ms, tf = run(n, hw, cin, 1024 if hw < 32 else 512, bn)
            row.append(f"bn{bn}: {ms * 1e3:7.1f} us {tf:6.0f} TF")
        print(f"batch {n:2d} {hw:2d}x{hw:<2d} C_in {cin:4d}: " + " | ".join(row), flush=True)
