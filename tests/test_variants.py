"""Model variants outside the BASELINE configs — the reference factory's own defaults (num_heads=4 with
num_head_channels=-1, resblock_updown=False -> Downsample / Upsample convs, script_util.py:44-62), ResBlocks without
use_scale_shift_norm (unet.py:253-255) — and the denoised_fn hook (gaussian_diffusion.py:262-265).
Fixtures: tests/golden/variants_golden.npz, produced by the REAL reference (oracle/make_golden_variants.py).
CPU tests pin the oracle and the state_dict layout; GPU tests check the CUDA path against the fixtures."""
import ctypes as C
import os

import numpy as np
import pytest
import torch as th
import torch.nn.functional as F

from guided_diffusion_clip_b200 import _lib as L
from guided_diffusion_clip_b200 import script_util as su
from oracle import golden_cfg as cfg
from oracle import oracle_diffusion as od
from oracle import oracle_models as om

TOL = 2e-2


@pytest.fixture(scope="module")
def V(golden_dir):
    return np.load(os.path.join(golden_dir, "variants_golden.npz"))


def _model(name):
    m = su.create_model(**cfg.VARIANT_KW[name])
    sd = om.make_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()},
                            cfg.VAR_SEED + sorted(cfg.VARIANT_KW).index(name))
    m.load_state_dict(sd, strict=True)
    return m, sd


def _tables(dk):
    return od.Tables(schedule=dk["noise_schedule"], steps=dk["steps"], respacing=dk.get("timestep_respacing", ""),
                     learn_sigma=dk.get("learn_sigma", False), sigma_small=dk.get("sigma_small", False),
                     predict_xstart=dk.get("predict_xstart", False))


# ------------------------------------------------------------------------------------------------ CPU
@pytest.mark.parametrize("name", sorted(cfg.VARIANT_KW))
def test_variant_state_dict_layout_and_oracle_match_reference(V, name):
    m, sd = _model(name)
    assert list(sd.keys()) == [str(k) for k in V[f"variant_{name}_keys"]]  # same keys in the same order
    x, t, y = cfg.variant_inputs(name)
    with th.no_grad():
        out = om.unet_forward(sd, x, t, y, **cfg.VARIANT_STRUCT[name])
    ref = th.from_numpy(V[f"variant_{name}_out"])
    err = float((out - ref).abs().max() / ref.abs().max())
    assert err < 2e-4, err


@pytest.mark.parametrize("name", cfg.DENOISED_CASES)
def test_denoised_fn_oracle_bit_exact(V, name):
    kw = cfg.STEP_CASES[name]
    tab = _tables(kw["diffusion"])
    xs, mo, g, i = cfg.step_inputs(name)
    th.manual_seed(cfg.STEP_NOISE_SEED)
    z = th.randn_like(xs)
    grad = g if kw["guided"] else None
    fn = cfg.denoised_fn_example
    r = (tab.ddim_sample(mo, xs, i, z, grad, eta=kw["eta"], denoised_fn=fn) if kw["ddim"]
         else tab.p_sample(mo, xs, i, z, grad, denoised_fn=fn))
    assert th.equal(r["sample"], th.from_numpy(V[f"denoised_{name}_sample"]))
    assert th.equal(r["pred_xstart"], th.from_numpy(V[f"denoised_{name}_x0"]))
    assert th.equal(tab.mean_variance(mo, xs, i, denoised_fn=fn)["mean"], th.from_numpy(V[f"denoised_{name}_mean"]))


@pytest.mark.parametrize("name", sorted(cfg.REVERSE_CASES))
def test_ddim_reverse_oracle_bit_exact(V, name):
    step, i = cfg.REVERSE_CASES[name]
    tab = _tables(cfg.STEP_CASES[step]["diffusion"])
    xs, mo, _, _ = cfg.step_inputs(step)
    r = tab.ddim_reverse_sample(mo, xs, i)
    assert th.equal(r["sample"], th.from_numpy(V[f"reverse_{name}_sample"]))
    assert th.equal(r["pred_xstart"], th.from_numpy(V[f"reverse_{name}_x0"]))


CLF_VARIANTS = {"plain": (cfg.CLF_PLAIN_KW, cfg.CLF_PLAIN_SEED), "convdown": (cfg.CLF_CONVDOWN_KW, cfg.CLF_CONVDOWN_SEED)}


@pytest.mark.parametrize("tag", sorted(CLF_VARIANTS))
def test_classifier_variant_oracle_matches_reference(V, tag):
    kw, seed = CLF_VARIANTS[tag]
    m = su.create_classifier(**kw)
    sd = om.make_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed)
    x, t, y = cfg.model_inputs()
    with th.no_grad():
        logits = om.classifier_forward(sd, x, t, **cfg.CLF_STRUCT)
    grad = om.classifier_guidance(sd, x, t, y, 1.0, **cfg.CLF_STRUCT)
    assert float((logits - th.from_numpy(V[f"clf_{tag}_logits"])).abs().max()) < 2e-4
    ref = th.from_numpy(V[f"clf_{tag}_grad"])
    assert float((grad - ref).abs().max() / ref.abs().max()) < 2e-4


def test_factory_defaults_build_a_model():
    """create_model_and_diffusion(**model_and_diffusion_defaults()) — the call every reference script makes — builds."""
    model, diffusion = su.create_model_and_diffusion(**su.model_and_diffusion_defaults())
    keys = list(model.state_dict().keys())
    assert "input_blocks.3.0.op.weight" in keys and "output_blocks.2.2.conv.weight" in keys
    assert model.state_dict()["input_blocks.1.0.emb_layers.1.weight"].shape == (256, 512)  # FiLM: 2 * 128 rows
    assert diffusion.num_timesteps == 1000


def test_new_entry_points_validate_arguments(lib):
    assert lib.gd_attention_fwd_hd(None, 0, None, 0, None, 1, 64, 1, 96, 0, None) != 0
    assert b"null" in lib.gd_last_error()
    buf = (C.c_char * 4096)()
    p = C.cast(buf, C.c_void_p)
    assert lib.gd_im2col3x3_s2_nhwc(p, 12, p, 108, 1, 4, 4, 12, None) != 0  # c % 8
    assert b"multiple of 8" in lib.gd_last_error()
    assert lib.gd_upsample2_nhwc(p, 8, p, 4, 1, 4, 4, 8, None) != 0        # ld_out < c
    assert lib.gd_add_emb_nhwc(p, 8, None, 8, 1, 16, 8, None) != 0


# ------------------------------------------------------------------------------------------------ GPU
def _rand(shape, seed, scale=1.0):
    g = th.Generator().manual_seed(seed)
    return (th.randn(shape, generator=g) * scale).cuda()


def _stream():
    return C.c_void_p(th.cuda.current_stream().cuda_stream)


def _rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(cfg.VARIANT_KW))
def test_variant_forward_matches_reference(lib, V, name):
    m, _ = _model(name)
    m = m.cuda().eval()
    x, t, y = cfg.variant_inputs(name)
    with th.no_grad():
        out = m(x.cuda(), t.cuda(), y.cuda() if y is not None else None)
    ref = th.from_numpy(V[f"variant_{name}_out"]).cuda()
    err = _rel(out, ref)
    print(f"variant {name}: rel err vs reference {err:.3e}")
    assert out.shape == ref.shape and err < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("d,t,heads,new_order", [(16, 64, 3, False), (32, 256, 2, True), (48, 256, 4, False),
                                                  (96, 100, 2, True), (128, 197, 2, False), (160, 64, 1, True),
                                                  (192, 320, 2, False), (256, 256, 1, True), (80, 65, 2, False),
                                                  (64, 72, 2, True), (112, 128, 1, False), (224, 64, 1, True)])
def test_attention_forward_any_head_width(lib, d, t, heads, new_order):
    n = 2
    qkv = _rand((n, 3 * heads * d, t), 50 + d).half().float()
    ref = om.qkv_attention(qkv, heads, new_order)  # [n, heads*d, t]
    qkv_b = qkv.permute(0, 2, 1).contiguous().half()
    out = th.zeros((n, t, heads * d), dtype=th.float16, device="cuda")
    lse = th.zeros((n, heads, t), dtype=th.float32, device="cuda")
    order = L.QKV_NEW if new_order else L.QKV_LEGACY
    L.check(lib.gd_attention_fwd_hd(qkv_b.data_ptr(), qkv_b.shape[-1], out.data_ptr(), out.shape[-1], lse.data_ptr(),
                                    n, t, heads, d, order, _stream()), "gd_attention_fwd_hd")
    th.cuda.synchronize()
    err = _rel(out.permute(0, 2, 1), ref)
    print(f"attention d={d} t={t} heads={heads} new={new_order}: rel err {err:.3e}")
    assert err < 4e-3
    # LSE of the scaled scores
    if new_order:
        q, k, _ = qkv.chunk(3, dim=1)
        q, k = q.reshape(n * heads, d, t), k.reshape(n * heads, d, t)
    else:
        q, k, _ = qkv.reshape(n * heads, 3 * d, t).split(d, dim=1)
    s = th.einsum("bct,bcs->bts", q, k) / d ** 0.5
    assert float((lse.reshape(n * heads, t) - th.logsumexp(s, dim=-1)).abs().max()) < 2e-3


@pytest.mark.gpu
@pytest.mark.parametrize("n,h,w,c,cout", [(2, 16, 16, 64, 64), (1, 32, 24, 128, 128), (3, 8, 8, 192, 192),
                                          (2, 15, 17, 64, 64)])
def test_strided_conv_and_upsample_conv(lib, n, h, w, c, cout):
    """Downsample.op (3x3 stride 2) = gd_im2col3x3_s2_nhwc + taps=1 GEMM; Upsample = gd_upsample2_nhwc + 3x3 conv."""
    from guided_diffusion_clip_b200.engine import pack_conv3x3
    from tests import gpu_helpers as H
    x = _rand((n, c, h, w), 61).half()
    wt = _rand((cout, c, 3, 3), 62, (c * 9) ** -0.5).half()
    b = _rand((cout,), 63, 0.1)
    xb = H.nhwc_half(x.float(), ld=c + 8, off=8)
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    cols = th.full((n, ho, wo, 9 * c), 7.0, dtype=th.float16, device="cuda")
    L.check(lib.gd_im2col3x3_s2_nhwc(xb.data_ptr() + 16, xb.shape[-1], cols.data_ptr(), 9 * c, n, h, w, c, _stream()))
    ref_cols = F.unfold(x.float(), 3, padding=1, stride=2).reshape(n, c, 9, ho, wo).permute(0, 3, 4, 2, 1)
    assert th.equal(cols.reshape(n, ho, wo, 9, c), ref_cols.half())  # a gather: bit-exact
    wp = pack_conv3x3(wt)
    out = H.conv_igemm(cols, 9 * c, 0, wp, b, cout, n, ho, wo, taps=1)
    ref = F.conv2d(x.float(), wt.float(), b, stride=2, padding=1)
    th.cuda.synchronize()
    err = _rel(out.permute(0, 3, 1, 2), ref)
    print(f"stride-2 conv {c}->{cout} {h}x{w}: rel err {err:.3e}")
    assert err < 2e-3
    up = th.zeros((n, 2 * h, 2 * w, c), dtype=th.float16, device="cuda")
    L.check(lib.gd_upsample2_nhwc(xb.data_ptr() + 16, xb.shape[-1], up.data_ptr(), c, n, h, w, c, _stream()))
    assert th.equal(up.permute(0, 3, 1, 2), F.interpolate(x, scale_factor=2, mode="nearest"))
    if (2 * h * 2 * w) % 8 == 0:
        out2 = H.conv_igemm(up, c, 0, wp, b, cout, n, 2 * h, 2 * w)
        ref2 = F.conv2d(F.interpolate(x.float(), scale_factor=2, mode="nearest"), wt.float(), b, padding=1)
        assert _rel(out2.permute(0, 3, 1, 2), ref2) < 2e-3


@pytest.mark.gpu
def test_add_embedding_in_place(lib):
    n, hw, c, ld = 3, 50, 72, 80
    x = _rand((n, hw, ld), 71).half()
    e = _rand((n, 200), 72)
    want = x.clone()
    want[..., 8:8 + c] = (x[..., 8:8 + c].float() + e[:, None, 40:40 + c]).half()
    L.check(lib.gd_add_emb_nhwc(x.data_ptr() + 16, ld, e.data_ptr() + 160, 200, n, hw, c, _stream()))
    th.cuda.synchronize()
    assert th.equal(x, want)  # one fp32 add, one rounding; untouched channels stay


@pytest.mark.gpu
@pytest.mark.parametrize("name", cfg.DENOISED_CASES)
def test_denoised_fn_matches_reference(lib, V, name):
    kw = cfg.STEP_CASES[name]
    d = su.create_gaussian_diffusion(**kw["diffusion"])
    xs, mo, g, i = (v.cuda() if isinstance(v, th.Tensor) else v for v in cfg.step_inputs(name))
    t = th.tensor([i] * xs.shape[0], device="cuda")
    th.manual_seed(cfg.STEP_NOISE_SEED)
    z = th.randn(xs.shape).cuda()
    model = lambda x, ts, **k: mo  # noqa: E731
    cond = (lambda x, ts, **k: g) if kw["guided"] else None  # noqa: E731
    r = d._sample_step(model, xs, t, True, cfg.denoised_fn_example, cond, {}, kw["ddim"], kw["eta"], noise=z)
    pmv = d.p_mean_variance(model, xs, t, denoised_fn=cfg.denoised_fn_example, model_kwargs={})
    for key, got in (("sample", r["sample"]), ("x0", r["pred_xstart"]), ("mean", pmv["mean"])):
        err = _rel(got, th.from_numpy(V[f"denoised_{name}_{key}"]).cuda())
        print(f"denoised_fn {name} {key}: rel err {err:.2e}")
        assert err < 2e-6
    # the public entry points accept the hook too
    out = (d.ddim_sample if kw["ddim"] else d.p_sample)(model, xs, t, denoised_fn=cfg.denoised_fn_example, cond_fn=cond,
                                                       model_kwargs={})
    assert out["sample"].shape == xs.shape and bool(th.isfinite(out["sample"]).all())


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(cfg.REVERSE_CASES))
def test_ddim_reverse_sample_matches_reference(lib, V, name):
    step, i = cfg.REVERSE_CASES[name]
    d = su.create_gaussian_diffusion(**cfg.STEP_CASES[step]["diffusion"])
    xs, mo, _, _ = (v.cuda() if isinstance(v, th.Tensor) else v for v in cfg.step_inputs(step))
    r = d.ddim_reverse_sample(lambda x, ts, **k: mo, xs, th.tensor([i] * xs.shape[0], device="cuda"), model_kwargs={})
    for key, got in (("sample", r["sample"]), ("x0", r["pred_xstart"])):
        err = _rel(got, th.from_numpy(V[f"reverse_{name}_{key}"]).cuda())
        print(f"ddim_reverse {name} {key}: rel err {err:.2e}")
        assert err < 2e-6
    with pytest.raises(AssertionError):
        d.ddim_reverse_sample(lambda x, ts, **k: mo, xs, th.tensor([i] * xs.shape[0], device="cuda"), eta=0.5)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ddpm_guided_mid", "ddim_guided_eta0", "ddpm_fixed_small"])
@pytest.mark.parametrize("hw", [(5, 5), (8, 6), (7, 4)])
def test_posterior_scalar_and_vector_paths_agree_with_oracle(lib, name, hw):
    """The fused update reads float4 when c*h*w % 4 == 0 and every pointer is 16-byte aligned, scalars otherwise: odd
    sizes (scalar kernel), aligned sizes (vector kernel) and a deliberately misaligned view of an aligned size all give
    the oracle's numbers; the two kernels are bit-identical to each other."""
    kw = cfg.STEP_CASES[name]
    d = su.create_gaussian_diffusion(**kw["diffusion"])
    tab = _tables(kw["diffusion"])
    learned = kw["diffusion"].get("learn_sigma", False)
    g = th.Generator().manual_seed(90 + hw[0])
    xs = th.randn((3, 3) + hw, generator=g)
    mo = th.randn((3, 6 if learned else 3) + hw, generator=g)
    gr = 0.3 * th.randn((3, 3) + hw, generator=g)
    z = th.randn((3, 3) + hw, generator=g)
    i = kw["index"]
    grad = gr if kw["guided"] else None
    ref = tab.ddim_sample(mo, xs, i, z, grad, eta=kw["eta"]) if kw["ddim"] else tab.p_sample(mo, xs, i, z, grad)
    t = th.full((3,), i, dtype=th.int64, device="cuda")

    def run(shift):
        def dev(v):  # shift = 1: place the tensor 4 bytes off a 16-byte boundary -> scalar kernel
            if v is None:
                return None
            buf = th.zeros(v.numel() + 4, device="cuda")
            return buf[shift:shift + v.numel()].view(v.shape).copy_(v.cuda())
        sample, x0 = dev(th.zeros_like(xs)), dev(th.zeros_like(xs))
        d._launch_posterior(x=dev(xs), t=t, model_out=dev(mo), grad=dev(grad), noise=dev(z), sample=sample,
                            pred_xstart=x0, ddim=kw["ddim"], eta=kw["eta"])
        th.cuda.synchronize()
        return sample.cpu(), x0.cpu()

    s0, p0 = run(0)
    s1, p1 = run(1)
    assert th.equal(s0, s1) and th.equal(p0, p1)
    assert _rel(s0, ref["sample"]) < 2e-6 and _rel(p0, ref["pred_xstart"]) < 2e-6


@pytest.mark.gpu
@pytest.mark.parametrize("tag", sorted(CLF_VARIANTS))
def test_classifier_variant_logits_and_gradient_match_reference(lib, V, tag):
    """classifier_use_scale_shift_norm=False / classifier_resblock_updown=False: forward and data-gradient (the
    reference's own closure through autograd, and the ClassifierGuidance fast path) against the real reference."""
    from guided_diffusion_clip_b200.sampler import ClassifierGuidance
    kw, seed = CLF_VARIANTS[tag]
    m = su.create_classifier(**kw)
    sd = om.make_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    x, t, y = (v.cuda() for v in cfg.model_inputs())
    with th.no_grad():
        logits = m(x, t)
    ref_l = th.from_numpy(V[f"clf_{tag}_logits"]).cuda()
    ref_g = th.from_numpy(V[f"clf_{tag}_grad"]).cuda()
    print(f"{tag} classifier logits rel err {_rel(logits, ref_l):.3e}")
    assert _rel(logits, ref_l) < TOL
    with th.enable_grad():
        x_in = x.detach().requires_grad_(True)
        sel = F.log_softmax(m(x_in, t), dim=-1)[range(len(x)), y.view(-1)]
        g1 = th.autograd.grad(sel.sum(), x_in)[0]
    g2 = ClassifierGuidance(m, 1.0)(x, t, y=y)
    print(f"{tag} classifier gradient rel err {_rel(g1, ref_g):.3e} (closure), {_rel(g2, ref_g):.3e} (guidance object)")
    assert _rel(g1, ref_g) < TOL and _rel(g2, ref_g) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("n,h,w,c", [(2, 16, 16, 64), (1, 9, 13, 128), (3, 8, 8, 192)])
def test_col2im_is_the_transpose_of_the_strided_gather(lib, n, h, w, c):
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    dcols = (_rand((n, ho, wo, 9 * c), 81) * 0.5).half()
    dx = th.full((n, h, w, c + 8), 3.0, dtype=th.float16, device="cuda")
    L.check(lib.gd_col2im3x3_s2_nhwc(dcols.data_ptr(), 9 * c, dx.data_ptr() + 16, c + 8, n, h, w, c, _stream()))
    th.cuda.synchronize()
    cols = dcols.float().reshape(n, ho * wo, 9, c).permute(0, 3, 2, 1).reshape(n, c * 9, ho * wo)  # F.fold layout
    ref = F.fold(cols, (h, w), 3, padding=1, stride=2)
    assert _rel(dx[..., 8:].permute(0, 3, 1, 2), ref) < 2e-3
    assert float((dx[..., :8] - 3.0).abs().max()) == 0.0


@pytest.mark.gpu
def test_unet_on_an_image_size_off_the_64_token_grid(lib):
    """48x48 input: attention over 12x12 = 144 tokens (64-wide heads go through the any-length kernel),
    ragged conv tiles everywhere; against the oracle on the same weights."""
    kw = dict(cfg.UNET_KW, image_size=48, attention_resolutions="12", channel_mult="1,2,4")
    m = su.create_model(**kw)
    sd = om.make_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 77)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    g = th.Generator().manual_seed(78)
    x = th.randn(3, 3, 48, 48, generator=g)
    t, y = th.tensor([5, 500, 999]), th.tensor([0, 10, 999])
    with th.no_grad():
        ref = om.unet_forward(sd, x, t, y, num_res_blocks=1, channel_mult_len=3, head_dim=64, new_order=True)
        out = m(x.cuda(), t.cuda(), y.cuda())
    err = _rel(out, ref.cuda())
    print(f"48x48 UNet vs oracle: rel err {err:.3e}")
    assert err < TOL
