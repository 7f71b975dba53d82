"""Microbenchmark of the attention forward kernels on BASELINE configs[1] shapes (UNet-256: 8 heads x T=1024,
16 heads x T=256, 16 heads x T=64; CLIP ViT-B/16: 12 heads x T=256 padded, 197 valid), batch 64: the tcgen05 / TMEM
kernel against the warp-level mma.sync kernel (gd_debug_set key 5).  FLOP = 4 * T_q * T_k * 64 per head."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch as th  # noqa: E402

from guided_diffusion_clip_b200 import _lib as L  # noqa: E402

CASES = [(64, 1024, 1024, 8, "legacy"), (64, 256, 256, 16, "legacy"), (64, 64, 64, 16, "legacy"),
         (32, 256, 197, 12, "new")]


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    evs = []
    for _ in range(reps):
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        evs.append((e0, e1))
    th.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    return ms[len(ms) // 2]


def main():
    lib = L.load()
    st = C.c_void_p(th.cuda.current_stream().cuda_stream)
    for n, t, tv, heads, order in CASES:
        c = heads * 64
        qkv = th.randn(n, t, 3 * c, device="cuda").half()
        out = th.empty(n, t, c, device="cuda", dtype=th.float16)
        o = L.QKV_LEGACY if order == "legacy" else L.QKV_NEW
        row = {"n": n, "t": t, "t_valid": tv, "heads": heads, "order": order}
        flop = 4.0 * n * heads * t * tv * 64
        for name, en in (("tcgen05", 1), ("mma_sync", 0)):
            lib.gd_debug_set(5, en)

            def fn():
                L.check(lib.gd_attention_fwd_masked(C.c_void_p(qkv.data_ptr()), 3 * c, C.c_void_p(out.data_ptr()), c,
                                                    None, n, t, tv, heads, o, st))
            ms = timeit(fn)
            row[name + "_ms"] = round(ms, 4)
            row[name + "_tflops"] = round(flop / ms / 1e9, 1)
        lib.gd_debug_set(5, 1)
        for k, share in ((0, "0/16"), (1, "4/16"), (2, "6/16"), (3, "8/16")):  # exponentials moved to the FMA pipe
            lib.gd_debug_set(5, 10 + k)
            ms = timeit(fn)
            row[f"poly_{share}_ms"] = round(ms, 4)
            row[f"poly_{share}_tflops"] = round(flop / ms / 1e9, 1)
        lib.gd_debug_set(5, 11)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
