"""Streaming-kernel bandwidth, board idle vs at its power cap (0.5 s of cuBLAS GEMMs right before each timing).
Separates DRAM-bound from SM-side-bound behaviour of GroupNorm-like kernels (gd_bw_probe in csrc/bw_probe.cu)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th  # noqa: E402

from guided_diffusion_clip_b200 import _lib as L  # noqa: E402

lib = L.load()
N = 64 * 256 * 256 * 256
x = th.randn(N, device="cuda").half()
y = th.empty_like(x)
a = th.randn((8192, 8192), device="cuda", dtype=th.float16)
st = C.c_void_p(th.cuda.current_stream().cuda_stream)


def run(structure, math, hot, reps=8):
    if hot:
        for _ in range(600):
            th.matmul(a, a)
    evs = []
    for _ in range(reps):
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record()
        if structure == -1000:
            y.copy_(x)
        else:
            L.check(lib.gd_bw_probe(structure, math, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), N * 2, st))
        e1.record()
        evs.append((e0, e1))
    th.cuda.synchronize()
    ms = sorted(p.elapsed_time(q) for p, q in evs)
    return 2 * N * 2 / (ms[len(ms) // 2] * 1e-3) / 1e9


for structure in (-1000, 0, -1, -2, -8, -32, 3):
    for math in ((0,) if structure == -1000 else (0, 2)):
        row = {"structure": {-1000: "torch copy_", 0: "flat one-shot 128thr x4"}.get(
                   structure, f"grid-stride {structure} CTA/SM" if structure > 0 else f"chunked 256thr, {-structure} rounds of 8"),
               "math": ["copy", "fma", "fma+silu"][math]}
        for hot in (False, True):
            row["hot_GBs" if hot else "idle_GBs"] = round(run(structure, math, hot))
        print(json.dumps(row), flush=True)

# the real GroupNorm apply (C=256 @256x256, batch 64) in the same harness
n, s, c = 64, 256, 256
gamma, beta = th.ones(c, device="cuda"), th.zeros(c, device="cuda")
film = th.zeros((n, 2 * c), device="cuda")
stats = th.zeros((n, 32, 2), device="cuda")
stats[:, :, 1] = 1.0
vp = lambda t: C.c_void_p(t.data_ptr())
for silu in (0, 1):
    row = {"structure": "gd_groupnorm_apply 64x256x256x256", "math": ["affine", "affine+silu"][silu]}
    for hot in (False, True):
        if hot:
            for _ in range(600):
                th.matmul(a, a)
        evs = []
        for _ in range(8):
            e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
            e0.record()
            L.check(lib.gd_groupnorm_apply(vp(x), c, vp(stats), vp(gamma), vp(beta), vp(film), 2 * c, vp(y), c, n, s, s, c,
                                           silu, L.GN_SAME, None, 0, st))
            e1.record()
            evs.append((e0, e1))
        th.cuda.synchronize()
        ms = sorted(p.elapsed_time(q) for p, q in evs)
        row["hot_GBs" if hot else "idle_GBs"] = round(2 * N * 2 / (ms[len(ms) // 2] * 1e-3) / 1e9)
        row["hot_all_ms" if hot else "idle_all_ms"] = [round(p.elapsed_time(q), 3) for p, q in evs]
    print(json.dumps(row), flush=True)
