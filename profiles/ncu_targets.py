"""The kernels bench.py's roofline fields refer to, launched once each after warm-up at the batch-64 shapes of the
step, for ONE `ncu --set full` capture of the round's FINAL build:

  python profiles/ncu_targets.py > gpurun_out/plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on --profile-from-start off \
      -o gpurun_out/ncu_targets_r02 python profiles/ncu_targets.py

Only the launches between cudaProfilerStart / Stop are captured; profiles/ncu_summarize.py turns the report into
profiles/ncu_summary_r02.json, which bench.py reads for `roofline.traffic` / `tensor_pipe_active_pct_ncu`."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th  # noqa: E402

from guided_diffusion_clip_b200 import _lib as L  # noqa: E402
from guided_diffusion_clip_b200 import script_util as su  # noqa: E402
from guided_diffusion_clip_b200.engine import pack_conv3x3, pack_conv3x3_bwd  # noqa: E402
from tests import gpu_helpers as H  # noqa: E402

B = int(os.environ.get("NCU_BATCH", "64"))
lib = L.load()
g = th.Generator().manual_seed(0)
dev = "cuda"


def conv_case(cin, cout, hw, gn, c1=0, res=False):
    x = th.randn((B, hw, hw, cin), generator=g).half().to(dev)
    w1 = (th.randn((cout, c1, 1, 1), generator=g) * c1 ** -0.5).to(dev) if c1 else None
    pack = pack_conv3x3((th.randn((cout, cin, 3, 3), generator=g) * (cin * 9) ** -0.5).to(dev), w1)
    bias = th.zeros(cout, device=dev)
    out = th.empty((B, hw, hw, cout), dtype=th.float16, device=dev)
    kw = {}
    if c1:  # the fused 1x1 skip_connection operand of an up ResBlock's second conv (unet.py:222, 256)
        kw.update(a1_buf=th.randn((B, hw, hw, c1), generator=g).half().to(dev), c1=c1)
    if res:  # identity residual (unet.py:256 with skip_connection = Identity)
        kw.update(res_buf=th.randn((B, hw, hw, cout), generator=g).half().to(dev), res_mode=L.RES_SAME)
    if gn:
        st = H.gn_stats(x, cin)
        film = (0.1 * th.randn((B, 2 * cin), generator=g)).to(dev)
        kw["gn"] = dict(mode=L.CONV_GN_SAME, silu=True,
                        coef=H.gn_coef(st, th.ones(cin, device=dev), th.zeros(cin, device=dev), film, B, cin))
    return lambda: H.conv_igemm(x, cin, 0, pack, bias, cout, B, hw, hw, out_buf=out, **kw)


def gn_bwd_case(c, hw):
    x = th.randn((B, hw, hw, c), generator=g).half().to(dev)
    dy = th.randn((B, hw, hw, c), generator=g).half().to(dev)
    add = th.randn((B, hw, hw, c), generator=g).half().to(dev)
    st = H.gn_stats(x, c)
    gamma, beta = th.ones(c, device=dev), th.zeros(c, device=dev)
    return lambda: H.gn_bwd(x, c, st, gamma, beta, dy, silu=True, add_buf=add)


def posterior_case(hw):
    d = su.create_gaussian_diffusion(steps=1000, learn_sigma=True, noise_schedule="linear", timestep_respacing="250")
    shape = (B, 3, hw, hw)
    x, gr, z = (th.randn(shape, device=dev) for _ in range(3))
    mo = th.randn((B, 6, hw, hw), device=dev)
    s, x0 = th.empty(shape, device=dev), th.empty(shape, device=dev)
    t = th.full((B,), 120, dtype=th.int64, device=dev)
    return lambda: d._launch_posterior(x=x, t=t, model_out=mo, grad=gr, noise=z, sample=s, pred_xstart=x0)


def attn_case(heads, tokens):
    qkv = th.randn((B, tokens, 3 * heads * 64), generator=g).half().to(dev)
    return lambda: H.attention_fwd(qkv, heads, L.QKV_LEGACY)


def gn_apply_pool_case(c, hw):
    x = th.randn((B, hw, hw, c), generator=g).half().to(dev)
    st = H.gn_stats(x, c)
    gamma, beta = th.ones(c, device=dev), th.zeros(c, device=dev)
    aux = th.empty((B, hw // 2, hw // 2, c), dtype=th.float16, device=dev)
    return lambda: H.gn_apply(x, c, st, gamma, beta, silu=True, mode=L.GN_AVGPOOL2, aux=aux)


TARGETS = [
    ("conv_gn_256_256_b64", conv_case(256, 256, 256, True)),     # the dominant kernel of the step (fused GroupNorm operand)
    ("conv_gn_128_128_b64", conv_case(128, 128, 256, True)),     # classifier forward, N = 128 tiles
    ("conv_plain_128_128_b64", conv_case(128, 128, 256, False)),  # classifier data-gradient convs (no fused operand)
    ("gn_bwd_128_b64", gn_bwd_case(128, 256)),
    ("gn_apply_pool_256_b64", gn_apply_pool_case(256, 256)),
    ("posterior_b64", posterior_case(256)),
    ("attn_fwd_tc_b64", attn_case(8, 1024)),
    ("conv_gn_512_256_b64", conv_case(512, 256, 256, True)),                 # K = 4608, the largest time share of the UNet
    ("conv_gn_256_256_skip512_b64", conv_case(256, 256, 256, True, c1=512)),  # K = 2816: fused 1x1-skip operand
    ("conv_gn_256_256_res_b64", conv_case(256, 256, 256, True, res=True)),    # identity residual staged through shared memory
]

if __name__ == "__main__":
    for _, fn in TARGETS:  # warm-up: function attributes, descriptor entry points, caches
        fn()
    th.cuda.synchronize()
    th.cuda.profiler.start()
    for name, fn in TARGETS:
        fn()
        th.cuda.synchronize()
    th.cuda.profiler.stop()
    print("targets:", [n for n, _ in TARGETS], "batch", B)
