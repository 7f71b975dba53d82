"""HBM copy bandwidth with the board idle vs at its power cap (0.5 s of cuBLAS fp16 GEMMs right before the timing):
tells how much of the GroupNorm kernels' in-step slowdown is the chip's own bandwidth under sw_power_cap."""
import torch as th

x = th.empty(64 * 256 * 256 * 256, dtype=th.float16, device="cuda")
y = th.empty_like(x)
a = th.randn((8192, 8192), device="cuda", dtype=th.float16)
b = th.randn((8192, 8192), device="cuda", dtype=th.float16)


def bw(hot):
    if hot:
        for _ in range(600):
            th.matmul(a, b)
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        y.copy_(x)
    e1.record()
    th.cuda.synchronize()
    return 2 * x.numel() * 2 / (e0.elapsed_time(e1) / 5) / 1e6


for _ in range(3):
    y.copy_(x)
th.cuda.synchronize()
for hot in (False, True, False, True):
    print({"hot": hot, "copy_GBs": round(bw(hot))}, flush=True)
