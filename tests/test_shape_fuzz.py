"""Geometry sweep: small UNets at image sizes, aspect ratios, batch sizes and level structures off the beaten path
(the BASELINE configs are all powers of two) against the CPU oracle on the same weights.  Exercises ragged conv tiles,
several-images-per-tile packing with half-empty groups, the fused-statistics fallback, pooled / upsampled GroupNorm on
odd tile counts, conv resampling, the first-conv kernel's fallback and attention at token counts off the 64 grid."""
import pytest
import torch as th

from guided_diffusion_clip_b200 import script_util as su
from oracle import golden_cfg as cfg
from oracle import oracle_models as om

pytestmark = pytest.mark.gpu
TOL = 2e-2

# (name, (H, W), batch, create_model overrides, oracle struct overrides)
CASES = [
    ("40sq_b5", (40, 40), 5, dict(image_size=40, attention_resolutions="10", channel_mult="1,2,4"), {}),
    ("24sq_b1", (24, 24), 1, dict(image_size=24, attention_resolutions="12,6", channel_mult="1,2,3"), {}),
    ("72sq_b2_2blocks", (72, 72), 2, dict(image_size=72, attention_resolutions="18", channel_mult="1,1,2",
                                         num_res_blocks=2), dict(num_res_blocks=2)),
    ("64x32_b3", (64, 32), 3, dict(image_size=64, attention_resolutions="16,8", channel_mult="1,2,2,4"), {}),
    ("32x96_b2", (32, 96), 2, dict(image_size=32, attention_resolutions="8", channel_mult="1,2,4"), {}),
    ("56sq_b7_convresample", (56, 56), 7, dict(image_size=56, attention_resolutions="14", channel_mult="1,2,2",
                                              resblock_updown=False), {}),
    ("16sq_b9_legacy", (16, 16), 9, dict(image_size=16, attention_resolutions="8,4", channel_mult="1,2,4",
                                        use_new_attention_order=False), dict(new_order=False)),
    ("20sq_b4_heads2", (20, 20), 4, dict(image_size=20, attention_resolutions="10,5", channel_mult="1,2,4",
                                        num_head_channels=-1, num_heads=2), dict(num_heads=2)),
    ("128x64_b1_wide", (128, 64), 1, dict(image_size=128, attention_resolutions="16", channel_mult="1,1,2,2",
                                         num_channels=128), {}),
]


@pytest.mark.parametrize("name,hw,batch,over,struct", CASES, ids=[c[0] for c in CASES])
def test_unet_geometry_sweep(lib, name, hw, batch, over, struct):
    kw = dict(cfg.UNET_KW)
    kw.update(over)
    m = su.create_model(**kw)
    sd = om.make_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 300 + len(name))
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    g = th.Generator().manual_seed(400 + len(name))
    x = th.randn((batch, 3) + hw, generator=g)
    t = th.randint(0, 1000, (batch,), generator=g)
    y = th.randint(0, 1000, (batch,), generator=g)
    okw = dict(num_res_blocks=kw["num_res_blocks"], channel_mult_len=len(kw["channel_mult"].split(",")), head_dim=64,
               new_order=kw["use_new_attention_order"])
    okw.update(struct)
    with th.no_grad():
        ref = om.unet_forward(sd, x, t, y, **okw)
        out = m(x.cuda(), t.cuda(), y.cuda()).cpu()
        # a second batch size through the same model: plans are per (batch, H, W) and must not interfere
        out1 = m(x[:1].cuda(), t[:1].cuda(), y[:1].cuda()).cpu()
    err = float((out - ref).abs().max() / ref.abs().max())
    err1 = float((out1 - ref[:1]).abs().max() / ref[:1].abs().max())
    print(f"geometry {name}: rel err {err:.3e} (batch {batch}), {err1:.3e} (first sample alone)")
    assert out.shape == ref.shape
    assert err < TOL and err1 < TOL


@pytest.mark.parametrize("hw,low_hw,batch", [((64, 64), (24, 24), 2), ((40, 40), (20, 12), 3), ((32, 64), (8, 16), 1),
                                             ((48, 48), (48, 48), 2), ((64, 64), (96, 80), 2)])
def test_superres_geometry_sweep(lib, hw, low_hw, batch):
    """SuperResModel with low_res at non-integer / non-square / identity / DOWN-scaling ratios: F.interpolate(low_res,
    (H, W), mode="bilinear") (unet.py:677-680) is gd_bilinear_upsample_nchw here."""
    kw = dict(cfg.SR_KW, large_size=64, small_size=16)
    m = su.sr_create_model(**kw)
    sd = om.make_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 500 + hw[0])
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    g = th.Generator().manual_seed(600 + low_hw[0])
    x = th.randn((batch, 3) + hw, generator=g)
    low = th.rand((batch, 3) + low_hw, generator=g) * 2 - 1
    t = th.randint(0, 1000, (batch,), generator=g)
    y = th.randint(0, 1000, (batch,), generator=g)
    with th.no_grad():
        ref = om.unet_forward(sd, x, t, y, low_res=low, **cfg.SR_STRUCT)
        out = m(x.cuda(), t.cuda(), low_res=low.cuda(), y=y.cuda()).cpu()
    err = float((out - ref).abs().max() / ref.abs().max())
    print(f"super-res {low_hw} -> {hw}: rel err {err:.3e}")
    assert err < TOL


@pytest.mark.parametrize("size,batch,over,levels", [
    (64, 5, {}, 4), (64, 1, dict(classifier_depth=2), 4), (128, 3, {}, 5),
    (128, 2, dict(classifier_attention_resolutions="32,16,8", classifier_width=128), 5)])
def test_classifier_guidance_geometry_sweep(lib, size, batch, over, levels):
    """Classifier logits and guidance gradient at odd batch sizes, depth 2 and the 5-level 128x128 structure
    (script_util.py:244-255 channel_mult table) against the oracle."""
    from guided_diffusion_clip_b200.sampler import ClassifierGuidance
    kw = dict(cfg.CLASSIFIER_KW, image_size=size)
    kw.update(over)
    m = su.create_classifier(**kw)
    sd = om.make_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 700 + size + batch)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    g = th.Generator().manual_seed(800 + batch)
    x = th.randn((batch, 3, size, size), generator=g)
    t = th.randint(0, 1000, (batch,), generator=g)
    y = th.randint(0, 1000, (batch,), generator=g)
    okw = dict(num_res_blocks=kw["classifier_depth"], channel_mult_len=levels, head_dim=64)
    with th.no_grad():
        ref_l = om.classifier_forward(sd, x, t, **okw)
        logits = m(x.cuda(), t.cuda()).cpu()
    ref_g = om.classifier_guidance(sd, x, t, y, 2.5, **okw)
    grad = ClassifierGuidance(m, 2.5)(x.cuda(), t.cuda(), y=y.cuda()).cpu()
    el = float((logits - ref_l).abs().max() / ref_l.abs().max())
    eg = float((grad - ref_g).abs().max() / ref_g.abs().max())
    print(f"classifier {size}x{size} batch {batch} {over}: logits rel err {el:.3e}, gradient rel err {eg:.3e}")
    assert el < TOL and eg < TOL


@pytest.mark.parametrize("hw,batch", [((64, 64), 3), ((96, 128), 1), ((200, 200), 2), ((50, 70), 5)])
def test_clip_guidance_geometry_sweep(lib, hw, batch):
    """CLIP guidance gradient for sampler resolutions equal to, above and below the encoder's input size, non-square
    and at odd batch sizes (bilinear resize + its transpose, padded-token attention) against the CLIP oracle."""
    from guided_diffusion_clip_b200 import clip as gclip
    from oracle import oracle_clip as oc
    enc = gclip.CLIPVisionEncoder(**cfg.CLIP_TINY)
    sd = cfg.clip_state_dict({k: tuple(v.shape) for k, v in enc.state_dict().items()})
    enc.load_state_dict(sd, strict=True)
    enc = enc.cuda().eval()
    g = th.Generator().manual_seed(900 + hw[0])
    x = th.randn((batch, 3) + hw, generator=g).clamp(-1, 1)
    txt = th.randn(batch, cfg.CLIP_TINY["projection_dim"], generator=g)
    txt = txt / txt.norm(dim=-1, keepdim=True)
    ckw = dict(heads=cfg.CLIP_TINY["num_attention_heads"], layers=cfg.CLIP_TINY["num_hidden_layers"],
               patch=cfg.CLIP_TINY["patch_size"], image_size=cfg.CLIP_TINY["image_size"])
    ref = oc.guidance(sd, x, txt, cfg.CLIP_SCALE, **ckw)
    got = gclip.CLIPGuidance(enc, txt.cuda(), cfg.CLIP_SCALE)(x.cuda(), None).cpu()
    err = float((got - ref).abs().max() / ref.abs().max())
    print(f"CLIP guidance {hw} batch {batch}: rel err {err:.3e}")
    assert err < TOL
