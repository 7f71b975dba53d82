"""GPU parity of the non-GEMM kernels through the C ABI: fused GroupNorm (+SiLU/FiLM/pool/upsample) forward and
data-gradient, fused attention forward/backward (both qkv orders), the posterior update against the reference
golden vectors, embedding pieces and the classifier pool head.  References are plain PyTorch fp32."""
import ctypes as C
import math
import os

import numpy as np
import pytest
import torch as th
import torch.nn.functional as F

from guided_diffusion_clip_b200 import _lib as L
from guided_diffusion_clip_b200 import script_util as su
from oracle import golden_cfg as cfg
from oracle import oracle_models as om
from tests import gpu_helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_reference():
    th.backends.cudnn.allow_tf32 = False
    th.backends.cuda.matmul.allow_tf32 = False


def _rand(shape, seed, scale=1.0):
    g = th.Generator().manual_seed(seed)
    return (th.randn(shape, generator=g) * scale).cuda()


def _h(x):
    return x.half().float()


# ------------------------------------------------------------------------------------------------ GroupNorm
@pytest.mark.parametrize("c,h,w", [(64, 16, 16), (192, 8, 8), (256, 64, 64), (96, 16, 16), (1536, 8, 8)])
@pytest.mark.parametrize("mode", [L.GN_SAME, L.GN_AVGPOOL2, L.GN_UPSAMPLE2])
def test_groupnorm_forward(lib, c, h, w, mode):
    n = 2
    x = _h(_rand((n, c, h, w), 1) * 1.5 + 0.3)
    gamma, beta = _rand((c,), 2) * 0.2 + 1, _rand((c,), 3) * 0.2
    film = _rand((n, 2 * c), 4) * 0.3
    y = F.group_norm(x, 32, gamma, beta, eps=1e-5) * (1 + film[:, :c, None, None]) + film[:, c:, None, None]
    y = F.silu(y)
    if mode == L.GN_AVGPOOL2:
        y = F.avg_pool2d(y, 2)
    elif mode == L.GN_UPSAMPLE2:
        y = F.interpolate(y, scale_factor=2, mode="nearest")
    xb = H.nhwc_half(x, ld=c + 32, off=32)
    st = H.gn_stats(xb, c, off=32)
    aux = th.zeros((n, h // 2, w // 2, c), dtype=th.float16, device="cuda") if mode == L.GN_AVGPOOL2 else None
    out = H.gn_apply(xb, c, st, gamma, beta, film=film, silu=True, mode=mode, off=32, aux=aux)
    th.cuda.synchronize()
    if aux is not None:  # side output: avgpool of the raw input (x_upd of a down ResBlock)
        assert H.rel_err(aux.permute(0, 3, 1, 2), F.avg_pool2d(x, 2)) < 2e-3
    ref_mean = x.reshape(n, 32, -1).mean(-1)
    assert float((st[..., 0] - ref_mean).abs().max()) < 1e-4
    err = H.rel_err(out.permute(0, 3, 1, 2), y)
    print(f"gn fwd c={c} {h}x{w} mode={mode}: rel err {err:.3e}")
    assert err < 3e-3


def test_groupnorm_plain_no_act(lib):
    n, c, h, w = 2, 128, 16, 16
    x = _h(_rand((n, c, h, w), 5))
    gamma, beta = _rand((c,), 6) * 0.2 + 1, _rand((c,), 7) * 0.2
    xb = H.nhwc_half(x)
    out = H.gn_apply(xb, c, H.gn_stats(xb, c), gamma, beta, silu=False)
    th.cuda.synchronize()
    assert H.rel_err(out.permute(0, 3, 1, 2), F.group_norm(x, 32, gamma, beta, eps=1e-5)) < 2e-3


@pytest.mark.parametrize("c,h,w", [(64, 16, 16), (192, 8, 8), (128, 32, 32)])
@pytest.mark.parametrize("mode,silu,use_film,add_mode", [
    (L.GN_SAME, True, True, None), (L.GN_SAME, False, False, L.GN_SAME), (L.GN_AVGPOOL2, True, False, L.GN_AVGPOOL2),
    (L.GN_SAME, True, False, L.GN_SAME), (L.GN_UPSAMPLE2, True, False, None)])
def test_groupnorm_backward(lib, c, h, w, mode, silu, use_film, add_mode):
    n = 2
    x = _h(_rand((n, c, h, w), 8) * 1.3 + 0.2).requires_grad_(True)
    gamma, beta = _rand((c,), 9) * 0.2 + 1, _rand((c,), 10) * 0.2
    film = _rand((n, 2 * c), 11) * 0.3 if use_film else None
    y = F.group_norm(x, 32, gamma, beta, eps=1e-5)
    if use_film:
        y = y * (1 + film[:, :c, None, None]) + film[:, c:, None, None]
    if silu:
        y = F.silu(y)
    if mode == L.GN_AVGPOOL2:
        y = F.avg_pool2d(y, 2)
    elif mode == L.GN_UPSAMPLE2:
        y = F.interpolate(y, scale_factor=2, mode="nearest")
    dy = _h(_rand(tuple(y.shape), 12))
    total = (y * dy).sum()
    add = None
    if add_mode == L.GN_SAME:
        add = _h(_rand((n, c, h, w), 13))
        total = total + (x * add).sum()
    elif add_mode == L.GN_AVGPOOL2:
        add = _h(_rand((n, c, h // 2, w // 2), 13))
        total = total + (F.avg_pool2d(x, 2) * add).sum()
    total.backward()
    xb = H.nhwc_half(x.detach())
    st = H.gn_stats(xb, c)
    dx = H.gn_bwd(xb, c, st, gamma, beta, H.nhwc_half(dy), film=film, silu=silu, mode=mode,
                  add_buf=H.nhwc_half(add) if add is not None else None, add_mode=add_mode or L.GN_SAME)
    th.cuda.synchronize()
    err = H.rel_err(dx.permute(0, 3, 1, 2), x.grad)
    print(f"gn bwd c={c} {h}x{w} mode={mode} silu={silu} film={use_film} add={add_mode}: rel err {err:.3e}")
    assert err < 5e-3


# ------------------------------------------------------------------------------------------------ attention
def _ref_attention(qkv_nct, heads, new_order):
    return om.qkv_attention(qkv_nct, heads, new_order)


@pytest.mark.parametrize("t,heads", [(64, 2), (256, 4), (1024, 3)])
@pytest.mark.parametrize("new_order", [False, True])
def test_attention_forward_backward(lib, t, heads, new_order):
    n = 2
    c3 = 3 * heads * 64
    qkv = _h(_rand((n, c3, t), 14)).requires_grad_(True)  # reference layout [N, 3*H*d, T]
    ref = _ref_attention(qkv, heads, new_order)           # [N, H*d, T]
    dout = _h(_rand(tuple(ref.shape), 15))
    ref.backward(dout)
    order = L.QKV_NEW if new_order else L.QKV_LEGACY
    qkv_b = qkv.detach().permute(0, 2, 1).contiguous().half()  # token-major [n, t, 3C]
    out, lse = H.attention_fwd(qkv_b, heads, order)
    th.cuda.synchronize()
    err = H.rel_err(out.permute(0, 2, 1), ref)
    print(f"attn fwd t={t} heads={heads} new={new_order}: rel err {err:.3e}")
    assert err < 4e-3
    dqkv = H.attention_bwd(qkv_b, out, dout.permute(0, 2, 1).contiguous().half(), lse, heads, order)
    th.cuda.synchronize()
    err = H.rel_err(dqkv.permute(0, 2, 1), qkv.grad)
    print(f"attn bwd t={t} heads={heads} new={new_order}: rel err {err:.3e}")
    assert err < 1e-2


# ------------------------------------------------------------------------------------------------ posterior
def _golden(golden_dir):
    return np.load(os.path.join(golden_dir, "models_golden.npz"))


@pytest.mark.parametrize("name", sorted(cfg.STEP_CASES))
def test_posterior_step_matches_reference_golden(lib, golden_dir, name):
    """One fused launch == the reference's p_sample / ddim_sample / p_mean_variance outputs (fixtures produced by the
    real reference, oracle/make_golden.py).  Tolerance 2e-6 relative: only expf/sqrtf ulps differ."""
    G = _golden(golden_dir)
    kw = cfg.STEP_CASES[name]
    d = su.create_gaussian_diffusion(**kw["diffusion"])
    xs, mo, g, i = cfg.step_inputs(name)
    th.manual_seed(cfg.STEP_NOISE_SEED)
    z = th.randn_like(xs)  # the reference drew exactly this with the CPU generator
    xs, mo, g, z = xs.cuda(), mo.cuda(), g.cuda(), z.cuda()
    t = th.full((xs.shape[0],), i, dtype=th.int64, device="cuda")
    sample, x0, mean, var, logvar = (th.empty_like(xs) for _ in range(5))
    d._launch_posterior(x=xs, t=t, model_out=mo, grad=g if kw["guided"] else None, noise=z, sample=sample,
                        pred_xstart=x0, mean=mean, var=var, logvar=logvar, ddim=kw["ddim"], eta=kw["eta"])
    th.cuda.synchronize()
    for key, got in (("sample", sample), ("x0", x0), ("mean", mean), ("var", var), ("logvar", logvar)):
        ref = th.from_numpy(G[f"step_{name}_{key}"]).cuda()
        err = H.rel_err(got, ref)
        print(f"posterior {name} {key}: rel err {err:.2e}")
        assert err < 2e-6, (name, key, err)


def test_p_mean_variance_and_public_step_api(lib, golden_dir):
    """The reference-facing methods (p_mean_variance, p_sample with arbitrary Python callables) agree with the golden."""
    G = _golden(golden_dir)
    name = "ddpm_guided_mid"
    kw = cfg.STEP_CASES[name]
    d = su.create_gaussian_diffusion(**kw["diffusion"])
    xs, mo, g, i = (v.cuda() if isinstance(v, th.Tensor) else v for v in cfg.step_inputs(name))
    t = th.tensor([i] * xs.shape[0], device="cuda")
    seen = {}

    def model(x, ts, **k):
        seen["t"] = ts.clone()
        return mo

    out = d.p_mean_variance(model, xs, t, model_kwargs={})
    assert set(out) == {"mean", "variance", "log_variance", "pred_xstart"}
    assert seen["t"].tolist() == [d.timestep_map[i]] * 2  # respace.py:123-128 maps index -> original timestep
    assert H.rel_err(out["mean"], th.from_numpy(G[f"step_{name}_mean"]).cuda()) < 2e-6
    th.manual_seed(cfg.STEP_NOISE_SEED)
    z = th.randn(xs.shape)
    th.cuda.manual_seed(0)
    r = d._sample_step(model, xs, t, True, None, lambda x, ts, **k: g, {}, False, 0.0, noise=z.cuda())
    assert H.rel_err(r["sample"], th.from_numpy(G[f"step_{name}_sample"]).cuda()) < 2e-6


# ------------------------------------------------------------------------------------------------ small pieces
def test_timestep_embedding_and_linear(lib):
    from guided_diffusion_clip_b200.nn import timestep_embedding
    t = th.tensor([0, 1, 37, 999], device="cuda")
    ref = om.timestep_embedding(t.cpu(), 256)
    got = timestep_embedding(t, 256)
    # arguments reach 999 where one fp32 ulp is 6e-5, so a 1-ulp difference in freq shows up at that size
    assert float((got.cpu() - ref).abs().max()) < 2e-4
    m, k, n = 5, 300, 77
    x, w, b, add = _rand((m, k), 16), _rand((n, k), 17, k ** -0.5), _rand((n,), 18), _rand((m, n), 19)
    y = th.empty((m, n), device="cuda")
    L.check(lib.gd_linear_f32(H.vp(x), k, H.vp(w), H.vp(b), H.vp(add), n, H.vp(y), n, m, k, n, 1, 1, H.stream()))
    th.cuda.synchronize()
    ref = F.silu(F.linear(F.silu(x), w, b) + add)
    assert H.rel_err(y, ref) < 1e-5


def test_uint8_pack_truncates_like_reference(lib):
    from guided_diffusion_clip_b200.dist_util import to_uint8_nhwc
    x = _rand((2, 3, 16, 16), 20, 0.8)
    x[0, 0, 0, :4] = th.tensor([-1.0, 1.0, 0.999, -0.0039], device="cuda")
    ref = ((x + 1) * 127.5).clamp(0, 255).to(th.uint8).permute(0, 2, 3, 1).contiguous()
    assert th.equal(to_uint8_nhwc(x), ref)  # integer output: bit-exact


def test_gather_buffer_output_stage_is_bit_exact(lib):
    """the pack kernel writes this rank's rows of the single gather buffer (classifier_sample.py:87-96);
    sample_sharded through it == the reference's expression, order and truncation (single process here; the
    world-size-2 order is covered on CPU by test_sharding_gloo.py and on GPUs by bench.py --gpus N)."""
    from guided_diffusion_clip_b200 import dist_util as du
    x = _rand((4, 3, 32, 32), 22, 0.9)
    y = th.tensor([1, 2, 3, 4], device="cuda")
    ref = ((x + 1) * 127.5).clamp(0, 255).to(th.uint8).permute(0, 2, 3, 1).contiguous()
    gb = du.GatherBuffer(4, 3, 32, 32, x.device)
    imgs, labs = gb.pack_and_gather(x, y)
    assert len(imgs) == 1 and th.equal(imgs[0], ref) and th.equal(labs[0], y)
    calls = []

    def sample_batch(classes):
        calls.append(classes.clone())
        return x

    arr, lab = du.sample_sharded(sample_batch, num_samples=6, batch_size=4, num_classes=1000, device=x.device)
    assert arr.shape == (6, 32, 32, 3) and arr.dtype.name == "uint8" and len(calls) == 2
    import numpy as np
    assert np.array_equal(arr, th.cat([ref, ref])[:6].cpu().numpy())
    assert np.array_equal(lab, th.cat(calls)[:6].cpu().numpy())


def test_logsoftmax_select_bwd(lib):
    n, k = 4, 1000
    logits = _rand((n, k), 21, 3.0).requires_grad_(True)
    y = th.tensor([0, 999, 5, 5], device="cuda")
    F.log_softmax(logits, -1)[range(n), y].sum().backward()
    d = th.empty((n, k), device="cuda")
    L.check(lib.gd_logsoftmax_select_bwd(H.vp(logits.detach()), H.vp(y), H.vp(d), n, k, C.c_float(2.5), H.stream()))
    th.cuda.synchronize()
    assert H.rel_err(d, 2.5 * logits.grad) < 1e-5


def test_attention_pool_head(lib):
    """AttentionPool2d forward and dX (unet.py:22-51) against autograd, T = 8*8 + 1 tokens, 4 heads."""
    n, c, s, n_out, heads = 2, 256, 8, 1000, 4
    hw = s * s
    hmap = _h(_rand((n, c, s, s), 22)).requires_grad_(True)
    pos = _rand((c, hw + 1), 23, c ** -0.5)
    wq, bq = _rand((3 * c, c), 24, c ** -0.5), _rand((3 * c,), 25, 0.1)
    wc, bc = _rand((n_out, c), 26, c ** -0.5), _rand((n_out,), 27, 0.1)
    tok = hmap.reshape(n, c, -1)
    tok = th.cat([tok.mean(-1, keepdim=True), tok], -1) + pos[None]
    a = om.qkv_attention(F.conv1d(tok, wq[..., None], bq), heads, True)
    ref = F.conv1d(a, wc[..., None], bc)[:, :, 0]
    dl = _rand((n, n_out), 28)
    ref.backward(dl)
    ws = th.empty(int(lib.gd_attnpool_ws_floats(n, hw + 1, c)), device="cuda")
    logits = th.empty((n, n_out), device="cuda")
    hb = H.nhwc_half(hmap.detach())
    L.check(lib.gd_attnpool_fwd(H.vp(hb), c, H.vp(pos), H.vp(wq), H.vp(bq), H.vp(wc), H.vp(bc), H.vp(logits), H.vp(ws),
                                n, hw, c, heads, n_out, H.stream()))
    th.cuda.synchronize()
    err = H.rel_err(logits, ref)
    print(f"attnpool fwd rel err {err:.3e}")
    assert err < 1e-4
    dh = th.zeros((n, s, s, c), dtype=th.float16, device="cuda")
    wq_t, wc_t = wq.t().contiguous(), wc.t().contiguous()  # must outlive the launch: the ABI only sees pointers
    L.check(lib.gd_attnpool_bwd(H.vp(dl), H.vp(wq_t), H.vp(wc_t), H.vp(ws), H.vp(dh), c,
                                n, hw, c, heads, n_out, C.c_float(1.0), H.stream()))
    th.cuda.synchronize()
    err = H.rel_err(dh.permute(0, 3, 1, 2), hmap.grad)
    print(f"attnpool bwd rel err {err:.3e}")
    assert err < 3e-3


@pytest.mark.parametrize("t,t_valid", [(128, 128), (256, 129), (256, 197), (256, 255), (384, 300), (1024, 1024)])
@pytest.mark.parametrize("new_order", [False, True])
def test_attention_tcgen05_forward_matches_fp32_and_mma_sync(lib, t, t_valid, new_order):
    """The tcgen05 / TMEM forward (sequence lengths that are multiples of 128) against a plain fp32 softmax(QK^T/8)V of
    the valid tokens, and against the warp-level mma.sync kernel of the same ABI call (gd_debug_set key 5 = 0), for both
    qkv channel orders, full and masked key ranges, with qkv / out rows strided inside wider buffers."""
    g = th.Generator().manual_seed(100 + t + t_valid)
    n, heads = 3, 3
    c = heads * 64
    ld_qkv, ld_out = 3 * c + 64, c + 8
    qkv = th.randn(n, t, ld_qkv, generator=g) * 1.5
    # later key tiles get larger and larger keys, so the running row maximum jumps by far more than the kernel's
    # lazy-rescale threshold (2^8) from tile to tile and the TMEM rescale of the output accumulator is exercised
    qkv[:, :, :3 * c] *= (1.0 + 2.0 * (th.arange(t) // 128).float() * (th.arange(t) % 3 == 0).float())[None, :, None]
    qkv = qkv.cuda().half()
    order = L.QKV_NEW if new_order else L.QKV_LEGACY
    outs, lses = [], []
    for enable in (1, 0):
        lib.gd_debug_set(5, enable)
        out = th.zeros(n, t, ld_out, device="cuda", dtype=th.float16)
        lse = th.zeros(n, heads, t, device="cuda")
        try:
            L.check(lib.gd_attention_fwd_masked(H.vp(qkv), ld_qkv, H.vp(out), ld_out, H.vp(lse), n, t, t_valid, heads, order,
                                                H.stream()))
            th.cuda.synchronize()
        finally:
            lib.gd_debug_set(5, 1)
        outs.append(out)
        lses.append(lse)
    x = qkv.float()[:, :t_valid, :3 * c]
    if new_order:
        q, k, v = (z.reshape(n, t_valid, heads, 64).transpose(1, 2) for z in x.chunk(3, dim=-1))
    else:
        q, k, v = (z.squeeze(3).transpose(1, 2) for z in x.reshape(n, t_valid, heads, 3, 64).split(1, dim=3))
    s = q @ k.transpose(-1, -2) / 8.0
    ref = (th.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(n, t_valid, c)
    ref_lse = th.logsumexp(s, dim=-1)
    e_tc = H.rel_err(outs[0][:, :t_valid, :c], ref)
    e_mma = H.rel_err(outs[1][:, :t_valid, :c], ref)
    print(f"attn tc t={t} valid={t_valid} new={new_order}: rel err tcgen05 {e_tc:.3e}, mma.sync {e_mma:.3e}")
    assert e_tc < 3e-3 and e_mma < 3e-3
    assert float((lses[0][:, :, :t_valid] - ref_lse).abs().max()) < 2e-3
    assert float(outs[0][:, :, c:].abs().max()) == 0.0  # nothing written outside the head columns
