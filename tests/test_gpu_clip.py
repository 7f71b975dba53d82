"""GPU parity of the CLIP ViT image-encoder guidance path against golden vectors produced by
transformers.CLIPVisionModelWithProjection (oracle/make_golden_clip.py) and against torch fp32 restatements of the
individual ops.  Tolerance: max-abs relative 2e-2 (fp16 storage vs the fp32 reference), as for the classifier path."""
import ctypes as C
import os

import numpy as np
import pytest
import torch as th
import torch.nn.functional as F

from guided_diffusion_clip_b200 import _lib as L
from guided_diffusion_clip_b200 import clip as gclip
from oracle import golden_cfg as cfg
from oracle import oracle_clip as oc
from tests import gpu_helpers as H

pytestmark = pytest.mark.gpu
TOL = 2e-2


def vp(t):
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(th.cuda.current_stream().cuda_stream)


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "clip_golden.npz"))


@pytest.fixture(scope="module")
def enc():
    m = gclip.CLIPVisionEncoder(**cfg.CLIP_TINY)
    sd = cfg.clip_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()})
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), sd


def test_layernorm_fwd_bwd_match_torch(lib):
    g = th.Generator().manual_seed(1)
    rows, c = 37, 768
    x = (th.randn(rows, c, generator=g) * 2 + 0.5).cuda().half()
    dy = th.randn(rows, c, generator=g).cuda().half()
    add = th.randn(rows, c, generator=g).cuda().half()
    gam = (1 + 0.1 * th.randn(c, generator=g)).cuda()
    bet = (0.1 * th.randn(c, generator=g)).cuda()
    out, dx = th.empty_like(x), th.empty_like(x)
    st = th.empty(rows, 2, device="cuda")
    L.check(lib.gd_layernorm_fwd(vp(x), c, vp(gam), vp(bet), C.c_float(1e-5), vp(out), c, vp(st), rows, c, stream()))
    L.check(lib.gd_layernorm_bwd(vp(x), c, vp(st), vp(gam), vp(dy), c, vp(add), c, vp(dx), c, rows, c, stream()))
    xr = x.float().requires_grad_(True)
    ref = F.layer_norm(xr, (c,), gam, bet, 1e-5)
    gref = th.autograd.grad(ref, xr, dy.float())[0] + add.float()
    assert H.rel_err(out.float(), ref.detach()) < 2e-3
    assert H.rel_err(dx.float(), gref) < 3e-3


def test_quickgelu_matches_torch(lib):
    g = th.Generator().manual_seed(2)
    rows, c = 19, 512
    x = (th.randn(rows, c, generator=g) * 3).cuda().half()
    dy = th.randn(rows, c, generator=g).cuda().half()
    out, dx = th.empty_like(x), th.empty_like(x)
    L.check(lib.gd_quickgelu_fwd(vp(x), c, vp(out), c, rows, c, stream()))
    L.check(lib.gd_quickgelu_bwd(vp(x), c, vp(dy), c, vp(dx), c, rows, c, stream()))
    xr = x.float().requires_grad_(True)
    ref = xr * th.sigmoid(1.702 * xr)
    gref = th.autograd.grad(ref, xr, dy.float())[0]
    assert H.rel_err(out.float(), ref.detach()) < 3e-3
    assert H.rel_err(dx.float(), gref) < 3e-3


@pytest.mark.parametrize("hin,size", [(80, 64), (64, 64), (50, 64), (256, 224)])
def test_clip_preprocess_and_its_transpose(lib, hin, size):
    """Fused (x+1)/2 -> bilinear (align_corners=False) -> CLIP mean/std -> 16x16 patches, and the exact transpose."""
    g = th.Generator().manual_seed(3)
    n, patch = 2, 16
    gsz = size // patch
    tp = (1 + gsz * gsz + 63) // 64 * 64
    k = 3 * patch * patch
    x = (th.rand(n, 3, hin, hin, generator=g) * 2 - 1).cuda()
    patches = th.full((n, tp, k), 7.0, device="cuda", dtype=th.float16)
    L.check(lib.gd_clip_preprocess_fwd(vp(x), vp(patches), k, n, hin, hin, size, patch, tp, stream()))
    xr = x.clone().requires_grad_(True)
    pix = oc.preprocess(xr.cpu(), size).cuda() if False else None
    mean = th.tensor(oc.CLIP_MEAN, device="cuda").view(1, 3, 1, 1)
    std = th.tensor(oc.CLIP_STD, device="cuda").view(1, 3, 1, 1)
    y = (xr + 1) / 2
    if hin != size:
        y = F.interpolate(y, size=(size, size), mode="bilinear", align_corners=False)
    y = (y - mean) / std
    ref = y.view(n, 3, gsz, patch, gsz, patch).permute(0, 2, 4, 1, 3, 5).reshape(n, gsz * gsz, k)
    got = patches.float()
    assert float(got[:, 0].abs().max()) == 0.0 and float(got[:, 1 + gsz * gsz:].abs().max()) == 0.0
    assert H.rel_err(got[:, 1:1 + gsz * gsz], ref.detach()) < 1.5e-3
    dp = th.zeros((n, tp, k), device="cuda", dtype=th.float16)
    dp[:, 1:1 + gsz * gsz] = th.randn(n, gsz * gsz, k, generator=g).cuda().half()
    dx = th.empty_like(x)
    L.check(lib.gd_clip_preprocess_bwd(vp(dp), k, vp(dx), n, hin, hin, size, patch, tp, C.c_float(1.0), stream()))
    gref = th.autograd.grad(ref, xr, dp[:, 1:1 + gsz * gsz].float())[0]
    assert H.rel_err(dx, gref) < 1e-4


@pytest.mark.parametrize("t_valid", [17, 64, 197])
def test_masked_attention_fwd_bwd(lib, t_valid):
    """Sequence padded to a multiple of 64: keys >= t_valid must not contribute, forward and backward."""
    g = th.Generator().manual_seed(4)
    n, heads = 2, 2
    t = (t_valid + 63) // 64 * 64
    c = heads * 64
    qkv = th.randn(n, t, 3 * c, generator=g).cuda().half()
    dout = th.randn(n, t, c, generator=g).cuda().half()
    dout[:, t_valid:] = 0
    out = th.empty(n, t, c, device="cuda", dtype=th.float16)
    lse = th.empty(n, heads, t, device="cuda")
    delta = th.empty(n, heads, t, device="cuda")
    dqkv = th.empty_like(qkv)
    L.check(lib.gd_attention_fwd_masked(vp(qkv), 3 * c, vp(out), c, vp(lse), n, t, t_valid, heads, L.QKV_NEW, stream()))
    L.check(lib.gd_attention_bwd_masked(vp(qkv), 3 * c, vp(out), c, vp(dout), c, vp(lse), vp(delta), vp(dqkv), 3 * c, n, t,
                                        t_valid, heads, L.QKV_NEW, stream()))
    x = qkv.float()[:, :t_valid].requires_grad_(True)
    q, k, v = (z.view(n, t_valid, heads, 64).transpose(1, 2) for z in x.chunk(3, dim=-1))
    a = th.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1) @ v
    ref = a.transpose(1, 2).reshape(n, t_valid, c)
    gref = th.autograd.grad(ref, x, dout.float()[:, :t_valid])[0]
    assert H.rel_err(out.float()[:, :t_valid], ref.detach()) < 3e-3
    assert H.rel_err(dqkv.float()[:, :t_valid], gref) < 5e-3
    if t_valid < t:  # padded keys receive exactly zero gradient
        assert float(dqkv.float()[:, t_valid:, c:].abs().max()) == 0.0


def test_clip_embedding_similarity_and_guidance_match_reference(lib, G, enc):
    m, sd = enc
    x, txt = cfg.clip_inputs()
    xc, tc = x.cuda(), txt.cuda()
    pooled = m.pooled(xc)
    e = pooled @ sd["visual_projection.weight"].cuda().t()
    ref = th.from_numpy(G["clip_embed"]).cuda()
    err = H.rel_err(e, ref)
    print(f"clip embedding rel err {err:.3e}")
    assert err < TOL
    plan = m.plan(x.shape[0], x.shape[2], x.shape[3], xc.device)
    sim = plan.similarity(xc, tc, cfg.CLIP_SCALE).clone()
    assert H.rel_err(sim, th.from_numpy(G["clip_sim"]).cuda()) < TOL
    grad = gclip.CLIPGuidance(m, tc, cfg.CLIP_SCALE)(xc, None)
    gref = th.from_numpy(G["clip_grad"]).cuda()
    gerr = H.rel_err(grad, gref)
    print(f"clip guidance gradient rel err {gerr:.3e}")
    assert gerr < TOL
    # and against the CPU oracle on the same weights (the golden pins the oracle, this pins the kernels to both)
    kw = dict(heads=cfg.CLIP_TINY["num_attention_heads"], layers=cfg.CLIP_TINY["num_hidden_layers"],
              patch=cfg.CLIP_TINY["patch_size"], image_size=cfg.CLIP_TINY["image_size"])
    go = oc.guidance(sd, x, txt, cfg.CLIP_SCALE, **kw)
    assert H.rel_err(grad.cpu(), go) < TOL


def test_clip_vit_b16_full_size_runs(lib):
    """BASELINE configs[2] shape: ViT-B/16 (197 tokens padded to 256), 256x256 input, batch 4: finite, batch
    independent gradient."""
    m = gclip.CLIPVisionEncoder().cuda().eval()
    g = th.Generator(device="cuda").manual_seed(5)
    x = th.rand((4, 3, 256, 256), generator=g, device="cuda") * 2 - 1
    txt = F.normalize(th.randn((1, 512), generator=g, device="cuda"), dim=-1)
    cond = gclip.CLIPGuidance(m, txt, 100.0)
    g4 = cond(x, None)
    g1 = cond(x[:1].clone(), None)
    assert th.isfinite(g4).all() and float(g4.abs().max()) > 0
    assert H.rel_err(g4[:1], g1) < 1e-3


def test_clip_guided_ddim_steps_match_oracle(lib, enc):
    """BASELINE configs[2] in miniature: unconditional ADM + CLIP guidance, the first 5 steps of a DDIM-50 chain
    through the public ddim_sample_loop_progressive API against the CPU oracle (UNet oracle pinned to the reference,
    CLIP oracle pinned to transformers, tables pinned bit-exact)."""
    from guided_diffusion_clip_b200 import script_util as su
    from oracle import oracle_diffusion as od
    from oracle import oracle_models as om
    m, clip_sd = enc
    kw = dict(cfg.UNET_KW, class_cond=False)
    unet = su.create_model(**kw)
    usd = om.make_state_dict({k: tuple(v.shape) for k, v in unet.state_dict().items()}, cfg.UNET_SEED + 100)
    unet.load_state_dict(usd, strict=True)
    unet.cuda().eval()
    diffusion = su.create_gaussian_diffusion(steps=1000, learn_sigma=True, noise_schedule="linear",
                                             timestep_respacing="ddim50")
    _, txt = cfg.clip_inputs()
    scale = 20.0 * cfg.CLIP_SCALE
    g = th.Generator().manual_seed(77)
    init = th.randn(2, 3, cfg.IMAGE, cfg.IMAGE, generator=g)
    steps = 5
    cond = gclip.CLIPGuidance(m, txt.cuda(), scale)
    gen = diffusion.ddim_sample_loop_progressive(unet, init.shape, noise=init.cuda(), cond_fn=cond, model_kwargs={},
                                                 device="cuda", eta=0.0)
    got = None
    for k, o in enumerate(gen):
        got = o
        if k + 1 == steps:
            break
    tab = od.Tables(schedule="linear", steps=1000, respacing="ddim50", learn_sigma=True)
    ckw = dict(heads=cfg.CLIP_TINY["num_attention_heads"], layers=cfg.CLIP_TINY["num_hidden_layers"],
               patch=cfg.CLIP_TINY["patch_size"], image_size=cfg.CLIP_TINY["image_size"])
    img = init.clone()
    with th.no_grad():
        for i in list(reversed(range(tab.T)))[:steps]:
            tt = th.full((2,), tab.timestep_map[i])
            mo = om.unet_forward(usd, img, tt, None, **cfg.UNET_STRUCT)
            gr = oc.guidance(clip_sd, img, txt, scale, **ckw)
            r = tab.ddim_sample(mo, img, i, th.zeros_like(img), gr, eta=0.0)
            img = r["sample"]
    err = H.rel_err(got["sample"].cpu(), img)
    print(f"CLIP-guided DDIM, {steps} steps: rel err {err:.3e}")
    assert err < TOL
