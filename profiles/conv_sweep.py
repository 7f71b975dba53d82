"""Microbenchmark of gd_conv_igemm on the conv problems of BASELINE configs[1], batch 8.
CUDA-event timing on the launching stream, inputs rotated over two buffers (each > L2 for the big layers).
Modes (gd_debug_set key 0): 0 = full kernel, 1 = mainloop only (epilogue = barriers), 2 = + TMEM loads.
Usage: python profiles/conv_sweep.py [--modes 0,1,2] [--bn 0]"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch as th  # noqa: E402

from guided_diffusion_clip_b200 import _lib as L  # noqa: E402

SHAPES = [
    # name, batch, H, Cin, Cout, taps, Cskip, res_mode
    ("256^2 256->256", 8, 256, 256, 256, 9, 0, L.RES_NONE),
    ("256^2 256->256 +res", 8, 256, 256, 256, 9, 0, L.RES_SAME),
    ("256^2 512->256", 8, 256, 512, 256, 9, 0, L.RES_NONE),
    ("256^2 256->256 +skip512", 8, 256, 256, 256, 9, 512, L.RES_NONE),
    ("128^2 256->256", 8, 128, 256, 256, 9, 0, L.RES_NONE),
    ("128^2 256->256 +poolres", 8, 128, 256, 256, 9, 0, L.RES_AVGPOOL2),
    ("64^2 512->512", 8, 64, 512, 512, 9, 0, L.RES_SAME),
    ("32^2 512->512", 8, 32, 512, 512, 9, 0, L.RES_SAME),
    ("16^2 1024->1024", 8, 16, 1024, 1024, 9, 0, L.RES_SAME),
    ("8^2 1024->1024", 8, 8, 1024, 1024, 9, 0, L.RES_SAME),
    ("8^2 2048->1024", 8, 8, 2048, 1024, 9, 0, L.RES_NONE),
    ("32^2 qkv 512->1536", 8, 32, 512, 1536, 1, 0, L.RES_NONE),
    ("256^2 128->128 (clf)", 8, 256, 128, 128, 9, 0, L.RES_NONE),
    ("256^2 128->128 +res (clf)", 8, 256, 128, 128, 9, 0, L.RES_SAME),
    ("256^2 1x1 64->256 (first conv)", 8, 256, 64, 256, 1, 0, L.RES_NONE),
    ("256^2 256->256 batch 64", 64, 256, 256, 256, 9, 0, L.RES_NONE),   # bench.py's roofline kernel
    ("256^2 128->128 batch 64 (clf)", 64, 256, 128, 128, 9, 0, L.RES_NONE),
    ("256^2 1x1 64->256 batch 64 (first conv)", 64, 256, 64, 256, 1, 0, L.RES_NONE),
    ("256^2 1x1 64->128 batch 64 (clf first conv)", 64, 256, 64, 128, 1, 0, L.RES_NONE),
]


WITH_STATS = False  # --stats: also request the fused GroupNorm partial statistics (as the UNet's convs do)


def run(shape, reps=8):
    name, n, h, cin, cout, taps, cskip, res_mode = shape
    lib = L.load()
    xs = [th.randn((n, h, h, cin), device="cuda", dtype=th.float16) for _ in range(2)]
    skip = th.randn((n, h, h, cskip), device="cuda", dtype=th.float16) if cskip else None
    rh = h * 2 if res_mode == L.RES_AVGPOOL2 else h
    res = th.randn((n, rh, rh, cout), device="cuda", dtype=th.float16) if res_mode != L.RES_NONE else None
    k = taps * cin + cskip
    w = (th.randn((cout, k), device="cuda") * k ** -0.5).half()
    bias = th.zeros(cout, device="cuda")
    out = th.empty((n, h, h, cout), device="cuda", dtype=th.float16)
    d = L.ConvDesc()
    d.c0, d.ld0, d.taps, d.n, d.h, d.w = cin, cin, taps, n, h, h
    if skip is not None:
        d.a1, d.c1, d.ld1 = skip.data_ptr(), cskip, cskip
    d.wpack, d.k_total, d.n_pad, d.bias, d.cout = w.data_ptr(), k, cout, bias.data_ptr(), cout
    if res is not None:
        d.res, d.ld_res, d.res_mode = res.data_ptr(), cout, res_mode
    d.out, d.ld_out, d.out_mode, d.out_scale = out.data_ptr(), cout, L.OUT_NHWC_F16, 1.0
    if WITH_STATS and cout % 64 == 0:
        stats = th.empty(((n * h * h + 127) // 128 * 4, cout // 4, 2), device="cuda")
        d.stats_out = stats.data_ptr()
    stream = C.c_void_p(th.cuda.current_stream().cuda_stream)
    evs = []
    for i in range(reps + 2):
        d.a0 = xs[i % 2].data_ptr()
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.gd_conv_igemm(C.byref(d), stream), "gd_conv_igemm")
        e1.record()
        evs.append((e0, e1))
    th.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs[2:])
    med = ms[len(ms) // 2]
    flops = 2.0 * n * h * h * cout * k
    return med, flops / (med * 1e-3) / 1e12


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--modes", default="0,1,2")
    ap.add_argument("--bn", default="0")
    ap.add_argument("--only", default="", help="substring filter on the shape name")
    ap.add_argument("--stats", action="store_true")
    args = ap.parse_args()
    WITH_STATS = args.stats
    if args.only:
        SHAPES[:] = [s for s in SHAPES if any(s[0] == f or (f.endswith("*") and s[0].startswith(f[:-1]))
                                              for f in args.only.split(","))]
    lib = L.load()
    results = []
    for bn in [int(v) for v in args.bn.split(",")]:
        lib.gd_debug_set(1, bn)
        for shape in SHAPES:
            row = {"shape": shape[0], "bn": bn}
            for mode in [int(m) for m in args.modes.split(",")]:
                lib.gd_debug_set(0, mode)
                try:
                    ms, tf = run(shape)
                    row[f"mode{mode}"] = {"ms": round(ms, 4), "tflops": round(tf, 1)}
                except Exception as e:  # noqa: BLE001
                    row[f"mode{mode}"] = str(e)[:80]
            lib.gd_debug_set(0, 0)
            print(json.dumps(row), flush=True)
            results.append(row)
