"""One 3x3 convolution with the GroupNorm (+FiLM+SiLU) operand transform fused (gd_conv_desc.gn_*) next to the
two-kernel path (gd_groupnorm_apply + plain conv), CUDA-event timed; the ncu target for the fused mainloop.

  python profiles/conv_gn_probe.py [batch] [hw] [cin] [cout] [reps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th  # noqa: E402

from guided_diffusion_clip_b200 import _lib as L  # noqa: E402
from guided_diffusion_clip_b200.engine import pack_conv3x3  # noqa: E402
from tests import gpu_helpers as H  # noqa: E402

n, hw, cin, cout, reps, c1 = (int(v) for v in (sys.argv[1:] + ["8", "256", "256", "256", "5", "0"][len(sys.argv) - 1:]))
g = th.Generator().manual_seed(0)
x = th.randn((n, hw, hw, cin), generator=g).half().cuda()
wt = (th.randn((cout, cin, 3, 3), generator=g) * (cin * 9) ** -0.5).cuda()
b = th.zeros(cout).cuda()
gamma, beta = th.ones(cin).cuda(), th.zeros(cin).cuda()
film = (0.1 * th.randn((n, 2 * cin), generator=g)).cuda()
st = H.gn_stats(x, cin)
w1 = (th.randn((cout, c1, 1, 1), generator=g) * c1 ** -0.5).cuda() if c1 else None
skip = th.randn((n, hw, hw, c1), generator=g).half().cuda() if c1 else None
pack = pack_conv3x3(wt, w1)
SK = dict(a1_buf=skip, c1=c1) if c1 else {}
out = th.empty((n, hw, hw, cout), dtype=th.float16, device="cuda")
gn = dict(mode=L.CONV_GN_SAME, silu=True, coef=H.gn_coef(st, gamma, beta, film, n, cin))


def timed(fn):
    fn()
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    th.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


fl = 2.0 * n * hw * hw * cout * (9 * cin + c1)
for na in (int(v) for v in os.environ.get("GN_NA", "").split(",") if v):
    L.load().gd_debug_set(7, na)
    t = timed(lambda: H.conv_igemm(x, cin, 0, pack, b, cout, n, hw, hw, out_buf=out, gn=gn, **SK))
    print(f"  activation ring depth {na}: fused {t:.3f} ms ({fl / t / 1e9:.0f} TF)")
L.load().gd_debug_set(7, 0)
t_f = timed(lambda: H.conv_igemm(x, cin, 0, pack, b, cout, n, hw, hw, out_buf=out, gn=gn, **SK))
if os.environ.get("QUAD_AB"):
    for q in (0, 1, 0, 1):
        L.load().gd_debug_set(8, q)
        normed_ = H.gn_apply(x, cin, st, gamma, beta, film=film, silu=True)
        tf_ = timed(lambda: H.conv_igemm(x, cin, 0, pack, b, cout, n, hw, hw, out_buf=out, gn=gn, **SK))
        tc_ = timed(lambda: H.conv_igemm(normed_, cin, 0, pack, b, cout, n, hw, hw, out_buf=out, **SK))
        print(f"  quad={q}: fused {tf_:.3f} ms ({fl / tf_ / 1e9:.0f} TF) | plain {tc_:.3f} ms ({fl / tc_ / 1e9:.0f} TF)")
    L.load().gd_debug_set(8, 1)
normed = H.gn_apply(x, cin, st, gamma, beta, film=film, silu=True)
t_c = timed(lambda: H.conv_igemm(normed, cin, 0, pack, b, cout, n, hw, hw, out_buf=out, **SK))
t_a = timed(lambda: H.gn_apply(x, cin, st, gamma, beta, film=film, silu=True))
print(f"n={n} {hw}x{hw} {cin}(+{c1} skip)->{cout}: fused {t_f:.3f} ms ({fl / t_f / 1e9:.0f} TF) | plain conv {t_c:.3f} ms "
      f"({fl / t_c / 1e9:.0f} TF) + gn_apply {t_a:.3f} ms = {t_c + t_a:.3f} ms")
