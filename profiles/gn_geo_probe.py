import ctypes as C, json, os, sys
sys.path.insert(0, '.')
import torch as th
from guided_diffusion_clip_b200 import _lib as L
lib = L.load()
n, s, c = 64, 256, int(os.environ.get("CH", "256"))
N = n * s * s * c
x = th.randn(N, device="cuda").half(); y = th.empty_like(x)
a = th.randn((8192, 8192), device="cuda", dtype=th.float16)
st = C.c_void_p(th.cuda.current_stream().cuda_stream)
gamma, beta = th.ones(c, device="cuda"), th.zeros(c, device="cuda")
film = th.zeros((n, 2 * c), device="cuda")
stats = th.zeros((n, 32, 2), device="cuda"); stats[:, :, 1] = 1.0
vp = lambda t: C.c_void_p(t.data_ptr())
for _ in range(600): th.matmul(a, a)
evs = []
for _ in range(10):
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    L.check(lib.gd_groupnorm_apply(vp(x), c, vp(stats), vp(gamma), vp(beta), vp(film), 2 * c, vp(y), c, n, s, s, c, 1, L.GN_SAME, None, 0, st))
    e1.record(); evs.append((e0, e1))
th.cuda.synchronize()
ms = sorted(p.elapsed_time(q) for p, q in evs)
print(json.dumps({"ppl": os.environ.get("GD_GN_PPL"), "thr": os.environ.get("GD_GN_THREADS"), "C": c, "GBs": round(2 * N * 2 / (ms[5] * 1e-3) / 1e9)}))
