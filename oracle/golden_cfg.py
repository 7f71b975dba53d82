"""ORACLE — TEST INFRASTRUCTURE ONLY.  Shared definition of the golden cases: configurations, seeds and
input generators used both by oracle/make_golden.py (which runs the reference) and by the tests."""
from __future__ import annotations

import torch as th

IMAGE = 64
UNET_SEED, CLF_SEED, INPUT_SEED, STEP_SEED, STEP_NOISE_SEED, TRAJ_SEED = 11, 12, 13, 14, 15, 16
CLF_SCALE = 1.0
TRAJ_BATCH = 2

# tiny class-conditional ADM: 64 -> 128 -> 192 -> 256 channels, attention at 16x16 and 8x8 (new order)
UNET_KW = dict(image_size=IMAGE, num_channels=64, num_res_blocks=1, channel_mult="", learn_sigma=True,
               class_cond=True, use_checkpoint=False, attention_resolutions="16,8", num_heads=4, num_head_channels=64,
               num_heads_upsample=-1, use_scale_shift_norm=True, dropout=0.0, resblock_updown=True, use_fp16=False,
               use_new_attention_order=True)
UNET_STRUCT = dict(num_res_blocks=1, channel_mult_len=4, head_dim=64, new_order=True)

# tiny classifier: width 64, depth 1, legacy attention at 16x16 and 8x8, attention pool over 8x8 (+1) tokens
CLASSIFIER_KW = dict(image_size=IMAGE, classifier_use_fp16=False, classifier_width=64, classifier_depth=1,
                     classifier_attention_resolutions="16,8", classifier_use_scale_shift_norm=True,
                     classifier_resblock_updown=True, classifier_pool="attention")
CLF_STRUCT = dict(num_res_blocks=1, channel_mult_len=4, head_dim=64)


def ref_unet_kwargs():
    """Constructor kwargs of the upstream-semantics reference unet.UNetModel for UNET_KW."""
    return dict(image_size=IMAGE, in_channels=3, model_channels=64, out_channels=6, num_res_blocks=1,
                attention_resolutions=(4, 8), dropout=0.0, channel_mult=(1, 2, 3, 4), num_classes=1000,
                use_checkpoint=False, use_fp16=False, num_heads=4, num_head_channels=64, num_heads_upsample=-1,
                use_scale_shift_norm=True, resblock_updown=True, use_new_attention_order=True)


def model_inputs():
    g = th.Generator().manual_seed(INPUT_SEED)
    x = th.randn(2, 3, IMAGE, IMAGE, generator=g)
    t = th.tensor([37, 961])
    y = th.tensor([3, 977])
    return x, t, y


def traj_labels():
    return th.tensor([5, 640])


SPACE_CASES = [(1000, "25"), (1000, "250"), (1000, "ddim25"), (1000, "ddim50"), (1000, "50"), (1000, "100"),
               (300, "10,15,20"), (1000, "1000"), (1000, "1"), (1000, "ddim1000"), (4000, "250"), (1000, "10"),
               (1000, "3,1,7"), (7, "7"), (1000, "ddim4"), (1000, "4"), (1000, "ddim7"), (1000, "3")]
SPACE_ERRORS = [(1000, "ddim999"), (10, "20"), (30, "5,20,5")]

_LIN250 = dict(steps=1000, learn_sigma=True, noise_schedule="linear", timestep_respacing="250")
DIFFUSION_CASES = {
    "linear_250_learned": _LIN250,
    "cosine_25_learned": dict(steps=1000, learn_sigma=True, noise_schedule="cosine", timestep_respacing="25"),
    "linear_ddim50_fixed": dict(steps=1000, learn_sigma=False, noise_schedule="linear", timestep_respacing="ddim50"),
    "linear_full_small": dict(steps=1000, learn_sigma=False, sigma_small=True, noise_schedule="linear"),
    "linear_10_rescaled": dict(steps=1000, learn_sigma=True, noise_schedule="linear", timestep_respacing="10",
                               rescale_timesteps=True),
    "cosine_4000_ddim25_xstart": dict(steps=4000, learn_sigma=True, noise_schedule="cosine",
                                      timestep_respacing="ddim25", predict_xstart=True),
}

# per-step arithmetic cases: (diffusion kwargs, ddim?, eta, guided?, step index)
STEP_CASES = {
    "ddpm_guided_mid": dict(diffusion=_LIN250, ddim=False, eta=0.0, guided=True, index=125),
    "ddpm_guided_t0": dict(diffusion=_LIN250, ddim=False, eta=0.0, guided=True, index=0),
    "ddpm_plain_last": dict(diffusion=_LIN250, ddim=False, eta=0.0, guided=False, index=249),
    "ddpm_fixed_large": dict(diffusion=DIFFUSION_CASES["linear_ddim50_fixed"], ddim=False, eta=0.0, guided=False,
                             index=17),
    "ddpm_fixed_small": dict(diffusion=DIFFUSION_CASES["linear_full_small"], ddim=False, eta=0.0, guided=True,
                             index=500),
    "ddim_guided_eta0": dict(diffusion=DIFFUSION_CASES["linear_ddim50_fixed"], ddim=True, eta=0.0, guided=True,
                             index=30),
    "ddim_plain_eta1": dict(diffusion=DIFFUSION_CASES["cosine_25_learned"], ddim=True, eta=1.0, guided=False,
                            index=12),
    "ddim_xstart_t0": dict(diffusion=DIFFUSION_CASES["cosine_4000_ddim25_xstart"], ddim=True, eta=0.5, guided=True,
                           index=0),
}


def step_inputs(name: str):
    kw = STEP_CASES[name]
    g = th.Generator().manual_seed(STEP_SEED + sorted(STEP_CASES).index(name))
    learn = kw["diffusion"].get("learn_sigma", False)
    xs = th.randn(2, 3, 8, 8, generator=g)
    mo = th.randn(2, 6 if learn else 3, 8, 8, generator=g)
    grad = 0.3 * th.randn(2, 3, 8, 8, generator=g)
    return xs, mo, grad, kw["index"]


_TR = dict(steps=1000, learn_sigma=True, noise_schedule="linear")
TRAJ_CASES = {
    "ddpm_guided_4": dict(diffusion=dict(_TR, timestep_respacing="4"), ddim=False, guided=True),
    "ddim_guided_4": dict(diffusion=dict(_TR, timestep_respacing="ddim4"), ddim=True, guided=True),
    "ddpm_plain_3": dict(diffusion=dict(_TR, timestep_respacing="3"), ddim=False, guided=False),
}


# ---- full-size configurations of BASELINE.json (layout hashes only; built on the meta device) --------
UNET256_KW = dict(image_size=256, num_channels=256, num_res_blocks=2, channel_mult="", learn_sigma=True,
                  class_cond=True, use_checkpoint=False, attention_resolutions="32,16,8", num_heads=4,
                  num_head_channels=64, num_heads_upsample=-1, use_scale_shift_norm=True, dropout=0.0,
                  resblock_updown=True, use_fp16=True, use_new_attention_order=False)
CLF256_KW = dict(image_size=256, classifier_use_fp16=False, classifier_width=128, classifier_depth=2,
                 classifier_attention_resolutions="32,16,8", classifier_use_scale_shift_norm=True,
                 classifier_resblock_updown=True, classifier_pool="attention")
SR512_KW = dict(large_size=512, small_size=128, num_channels=192, num_res_blocks=2, learn_sigma=True,
                class_cond=True, use_checkpoint=False, attention_resolutions="32,16", num_heads=4,
                num_head_channels=64, num_heads_upsample=-1, use_scale_shift_norm=True, dropout=0.0,
                resblock_updown=True, use_fp16=True)


def ref_unet256_kwargs():
    return dict(image_size=256, in_channels=3, model_channels=256, out_channels=6, num_res_blocks=2,
                attention_resolutions=(8, 16, 32), dropout=0.0, channel_mult=(1, 1, 2, 2, 4, 4), num_classes=1000,
                use_checkpoint=False, use_fp16=True, num_heads=4, num_head_channels=64, num_heads_upsample=-1,
                use_scale_shift_norm=True, resblock_updown=True, use_new_attention_order=False)


def ref_sr512_kwargs():
    return dict(image_size=512, in_channels=3, model_channels=192, out_channels=6, num_res_blocks=2,
                attention_resolutions=(16, 32), dropout=0.0, channel_mult=(1, 1, 2, 2, 4, 4), num_classes=1000,
                use_checkpoint=False, use_fp16=True, num_heads=4, num_head_channels=64, num_heads_upsample=-1,
                use_scale_shift_norm=True, resblock_updown=True)


# well-conditioned 10-step prefixes of the real chains (north_star: "short 10-step trajectories"): the first 10
# reverse steps of the 250-step ancestral chain and of the 50-step DDIM chain, classifier-guided
TRAJ10_STEPS = 10
TRAJ10_CASES = {
    "ddpm250_guided": dict(diffusion=dict(_TR, timestep_respacing="250"), ddim=False, guided=True),
    "ddim50_guided": dict(diffusion=dict(_TR, timestep_respacing="ddim50"), ddim=True, guided=True),
    "ddpm250_plain": dict(diffusion=dict(_TR, timestep_respacing="250"), ddim=False, guided=False),
}


def traj10_noise():
    """The reference's CPU-generator draws: randn(shape) then one randn_like per step (gaussian_diffusion.py:516,430)."""
    th.manual_seed(TRAJ_SEED)
    init = th.randn(TRAJ_BATCH, 3, IMAGE, IMAGE)
    return init, [th.randn(TRAJ_BATCH, 3, IMAGE, IMAGE) for _ in range(TRAJ10_STEPS)]


# ---- §8f rows already built: super-resolution input path and the fork's clip_feat conditioning ------------
SR_SEED, FEAT_SEED = 21, 22
SR_KW = dict(large_size=64, small_size=16, num_channels=64, num_res_blocks=1, learn_sigma=True, class_cond=True,
             use_checkpoint=False, attention_resolutions="16,8", num_heads=4, num_head_channels=64,
             num_heads_upsample=-1, use_scale_shift_norm=True, dropout=0.0, resblock_updown=True, use_fp16=False)
SR_STRUCT = dict(num_res_blocks=1, channel_mult_len=4, head_dim=64, new_order=False)
FEAT_KW = dict(UNET_KW, use_new_attention_order=False)  # create_model(..., conditioning="clip_feat")
FEAT_STRUCT = dict(num_res_blocks=1, channel_mult_len=4, head_dim=64, new_order=False)


def ref_sr_kwargs():
    return dict(image_size=64, in_channels=3, model_channels=64, out_channels=6, num_res_blocks=1,
                attention_resolutions=(4, 8), dropout=0.0, channel_mult=(1, 2, 3, 4), num_classes=1000,
                use_checkpoint=False, use_fp16=False, num_heads=4, num_head_channels=64, num_heads_upsample=-1,
                use_scale_shift_norm=True, resblock_updown=True)


def ref_feat_kwargs():
    kw = ref_unet_kwargs()
    kw.update(num_classes=512, use_new_attention_order=False)
    return kw


def sr_inputs():
    g = th.Generator().manual_seed(INPUT_SEED + 1)
    x = th.randn(2, 3, IMAGE, IMAGE, generator=g)
    low = th.rand(2, 3, 16, 16, generator=g) * 2 - 1
    return x, th.tensor([500, 3]), th.tensor([1, 999]), low


def feat_inputs():
    g = th.Generator().manual_seed(INPUT_SEED + 2)
    x = th.randn(2, 3, IMAGE, IMAGE, generator=g)
    feat = th.randn(2, 1, 512, generator=g)
    return x, th.tensor([10, 720]), feat / feat.norm(dim=-1, keepdim=True)


# ---- §8f row 2: CLIP ViT image-encoder guidance (spec: SURVEY §8c; stand-in reference: transformers CLIP) ----------
CLIP_SEED = 31
CLIP_TINY = dict(hidden_size=128, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2, image_size=64,
                 patch_size=16, projection_dim=64)
CLIP_INPUT = 80          # sampler resolution fed to the guidance: 80 -> 64 exercises the bilinear resize (scale 1.25)
CLIP_SCALE = 10.0


def clip_state_dict(shapes):
    """Seeded weights in the transformers key layout: LayerNorm weights around 1, biases / class token small,
    matrices ~ N(0, 1/fan_in), positions ~ N(0, 0.02)."""
    g = th.Generator().manual_seed(CLIP_SEED)
    sd = {}
    for name, shape in shapes.items():
        shape = tuple(shape)
        if "norm" in name and name.endswith("weight"):
            v = 1.0 + 0.1 * th.randn(shape, generator=g)
        elif name.endswith("position_embedding.weight"):
            v = 0.02 * th.randn(shape, generator=g)
        elif len(shape) == 1:
            v = 0.1 * th.randn(shape, generator=g)
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            v = th.randn(shape, generator=g) / fan_in ** 0.5
        sd[name] = v
    return sd


def clip_inputs():
    g = th.Generator().manual_seed(INPUT_SEED + 3)
    x = th.randn(2, 3, CLIP_INPUT, CLIP_INPUT, generator=g).clamp(-1, 1)
    txt = th.randn(2, CLIP_TINY["projection_dim"], generator=g)
    return x, txt / txt.norm(dim=-1, keepdim=True)


# ---- model variants outside the BASELINE configs: the reference factory's own defaults (script_util.py:44-62:
# num_heads=4 / num_head_channels=-1, resblock_updown=False -> Downsample / Upsample with conv_resample) and
# improved-diffusion style blocks (use_scale_shift_norm=False); plus the denoised_fn hook of p_mean_variance -------
VAR_SEED = 41
VARIANT_KW = {
    # additive embedding, conv resampling, heads of 48 (192/4) and 64 (256/4) channels, legacy qkv order
    "plain": dict(image_size=IMAGE, num_channels=64, num_res_blocks=1, channel_mult="", learn_sigma=True,
                  class_cond=True, use_checkpoint=False, attention_resolutions="16,8", num_heads=4,
                  num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=False, dropout=0.0,
                  resblock_updown=False, use_fp16=False, use_new_attention_order=False),
    # model_and_diffusion_defaults() (script_util.py:44-62) with class_cond: 128 channels, 2 blocks per level,
    # heads of 96 (384/4) and 128 (512/4) channels, FiLM, conv resampling
    "defaults": dict(image_size=IMAGE, num_channels=128, num_res_blocks=2, channel_mult="", learn_sigma=False,
                     class_cond=True, use_checkpoint=False, attention_resolutions="16,8", num_heads=4,
                     num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=True, dropout=0.0,
                     resblock_updown=False, use_fp16=False, use_new_attention_order=False),
    # new qkv order with two wide heads upstream and one 256-wide head at the top (num_heads_upsample differs)
    "wide_heads": dict(image_size=IMAGE, num_channels=64, num_res_blocks=1, channel_mult="", learn_sigma=True,
                       class_cond=False, use_checkpoint=False, attention_resolutions="8", num_heads=1,
                       num_head_channels=-1, num_heads_upsample=2, use_scale_shift_norm=True, dropout=0.0,
                       resblock_updown=True, use_fp16=False, use_new_attention_order=True),
}
VARIANT_STRUCT = {
    "plain": dict(num_res_blocks=1, channel_mult_len=4, num_heads=4, new_order=False),
    "defaults": dict(num_res_blocks=2, channel_mult_len=4, num_heads=4, new_order=False),
    "wide_heads": dict(num_res_blocks=1, channel_mult_len=4, num_heads=1, num_heads_upsample=2, new_order=True),
}


def ref_variant_kwargs(name: str):
    """Constructor kwargs of the reference unet.UNetModel that script_util.create_model derives from VARIANT_KW[name]
    (script_util.py:130-167; upstream nn.Embedding label semantics, as for ref_unet_kwargs)."""
    kw = VARIANT_KW[name]
    ds = tuple(IMAGE // int(r) for r in kw["attention_resolutions"].split(","))
    return dict(image_size=IMAGE, in_channels=3, model_channels=kw["num_channels"],
                out_channels=6 if kw["learn_sigma"] else 3, num_res_blocks=kw["num_res_blocks"],
                attention_resolutions=ds, dropout=0.0, channel_mult=(1, 2, 3, 4),
                num_classes=1000 if kw["class_cond"] else None, use_checkpoint=False, use_fp16=False,
                num_heads=kw["num_heads"], num_head_channels=kw["num_head_channels"],
                num_heads_upsample=kw["num_heads_upsample"], use_scale_shift_norm=kw["use_scale_shift_norm"],
                resblock_updown=kw["resblock_updown"], use_new_attention_order=kw["use_new_attention_order"])


def variant_inputs(name: str):
    g = th.Generator().manual_seed(INPUT_SEED + 10 + sorted(VARIANT_KW).index(name))
    x = th.randn(2, 3, IMAGE, IMAGE, generator=g)
    return x, th.tensor([250, 7]), (th.tensor([17, 800]) if VARIANT_KW[name]["class_cond"] else None)


def denoised_fn_example(x0: th.Tensor) -> th.Tensor:
    """A non-trivial denoised_fn (gaussian_diffusion.py:262-265 applies it before the clamp)."""
    return 0.8 * x0 + 0.1 * x0.flip(-1)


DENOISED_CASES = ["ddpm_guided_mid", "ddim_guided_eta0", "ddim_xstart_t0"]

# ddim_reverse_sample (gaussian_diffusion.py:596-632): (STEP_CASES entry whose diffusion / inputs are reused, step index)
REVERSE_CASES = {"rev_eps_mid": ("ddim_guided_eta0", 30), "rev_eps_last": ("ddim_guided_eta0", 49),
                 "rev_xstart_t0": ("ddim_xstart_t0", 0), "rev_learned": ("ddim_plain_eta1", 12)}


# ---- BASELINE configs[0] at FULL size: 64x64 class-conditional ADM (README.md:49 flags), unguided p_sample_loop,
# timestep_respacing 25, batch 4 — the reference's own CPU-runnable case ------------------------------------------
C1_SEED, C1_NOISE_SEED, C1_BATCH, C1_STEPS = 51, 52, 4, 25
C1_KW = dict(image_size=64, num_channels=192, num_res_blocks=3, channel_mult="", learn_sigma=True, class_cond=True,
             use_checkpoint=False, attention_resolutions="32,16,8", num_heads=4, num_head_channels=64,
             num_heads_upsample=-1, use_scale_shift_norm=True, dropout=0.0, resblock_updown=True, use_fp16=False,
             use_new_attention_order=True)
C1_DIFFUSION = dict(steps=1000, learn_sigma=True, noise_schedule="cosine", timestep_respacing="25")
C1_STRUCT = dict(num_res_blocks=3, channel_mult_len=4, head_dim=64, new_order=True)
C1_CHECKPOINTS = (10, 25)  # number of reverse steps after which the sample is recorded


def ref_c1_kwargs():
    return dict(image_size=64, in_channels=3, model_channels=192, out_channels=6, num_res_blocks=3,
                attention_resolutions=(2, 4, 8), dropout=0.0, channel_mult=(1, 2, 3, 4), num_classes=1000,
                use_checkpoint=False, use_fp16=False, num_heads=4, num_head_channels=64, num_heads_upsample=-1,
                use_scale_shift_norm=True, resblock_updown=True, use_new_attention_order=True)


def c1_labels():
    return th.tensor([1, 250, 500, 999])


def c1_noise():
    """The CPU-generator draws of the reference loop: randn(shape), then one randn_like per step."""
    th.manual_seed(C1_NOISE_SEED)
    shape = (C1_BATCH, 3, 64, 64)
    return th.randn(*shape), [th.randn(*shape) for _ in range(C1_STEPS)]

# classifier without use_scale_shift_norm (classifier_defaults() flag; additive embedding in every ResBlock)
CLF_PLAIN_SEED = 45
CLF_PLAIN_KW = dict(CLASSIFIER_KW, classifier_use_scale_shift_norm=False)
CLF_CONVDOWN_SEED = 46
CLF_CONVDOWN_KW = dict(CLASSIFIER_KW, classifier_resblock_updown=False)  # Downsample.op = 3x3 stride-2 conv


# ---- FULL-SIZE one-step fixtures of BASELINE configs[1..4] (oracle/make_golden_fullsize.py -> tests/golden/fullsize_*.npz)
# One guided step at batch 1 through the REAL reference in fp32, on make_state_dict weights; the GPU tests rebuild the
# same weights from the seed and compare eps|v, the guidance gradient and x_{t-1} at the real widths.
FS_SEED = 61  # weights: FS_SEED + k; inputs: FS_SEED + 10 + k; step noise: FS_SEED + 20 + k
UNET256U_KW = dict(UNET256_KW, class_cond=False)
UNET512_KW = dict(UNET256_KW, image_size=512, use_fp16=False)   # channel_mult (0.5,1,1,2,2,4,4), script_util.py:149-151
CLF512_KW = dict(CLF256_KW, image_size=512)
CLIP_B16 = dict(hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12, image_size=224,
                patch_size=16, projection_dim=512)
FULLSIZE_CASES = {
    # configs[1]: 256x256 class-cond ADM + classifier-256 guidance (scale 1.0), ancestral step 120 of 250
    "cfg2": dict(image=256, k=0, diffusion=dict(_TR, timestep_respacing="250"), ddim=False, index=120, label=207,
                 scale=1.0),
    # configs[2]: 256x256 unconditional ADM + CLIP ViT-B/16 guidance, DDIM step 30 of 50
    "cfg3": dict(image=256, k=1, diffusion=dict(_TR, timestep_respacing="ddim50"), ddim=True, index=30, label=None,
                 scale=100.0),
    # configs[3]: 128 -> 512 upsampler, ancestral step 100 of 250, unguided
    "cfg4": dict(image=512, k=2, diffusion=dict(_TR, timestep_respacing="250"), ddim=False, index=100, label=77,
                 scale=0.0),
    # configs[4]: 512x512 class-cond ADM (use_fp16 False) + classifier-512 guidance (scale 4.0), DDIM step 12 of 25
    "cfg5": dict(image=512, k=3, diffusion=dict(_TR, timestep_respacing="ddim25"), ddim=True, index=12, label=931,
                 scale=4.0),
}


def ref_unet512_kwargs():
    return dict(ref_unet256_kwargs(), image_size=512, attention_resolutions=(16, 32, 64),
                channel_mult=(0.5, 1, 1, 2, 2, 4, 4), use_fp16=False)


def fullsize_inputs(name: str):
    """x_t (unit normal, as at a mid-chain step), the low-res conditioning of the upsampler and the CLIP text vector."""
    c = FULLSIZE_CASES[name]
    g = th.Generator().manual_seed(FS_SEED + 10 + c["k"])
    x = th.randn(1, 3, c["image"], c["image"], generator=g)
    low = th.rand(1, 3, 128, 128, generator=g) * 2 - 1
    txt = th.randn(1, 512, generator=g)
    return x, low, txt / txt.norm(dim=-1, keepdim=True)


def fullsize_noise(name: str):
    """The CPU-generator draw of the reference's p_sample / ddim_sample (th.randn_like, gaussian_diffusion.py:430,585)."""
    c = FULLSIZE_CASES[name]
    th.manual_seed(FS_SEED + 20 + c["k"])
    return th.randn(1, 3, c["image"], c["image"])


FS_TRAJ_STEPS = 10  # first reverse steps of the 250-step guided chain of configs[1] at FULL size (batch 1)


def fullsize_traj_noise():
    """The CPU-generator draws of the reference loop for the full-size trajectory: randn(shape), then one randn_like per
    step (gaussian_diffusion.py:516, 430)."""
    th.manual_seed(FS_SEED + 30)
    shape = (1, 3, 256, 256)
    return th.randn(*shape), [th.randn(*shape) for _ in range(FS_TRAJ_STEPS)]


def fs_pack(a):
    """fixture storage: fp16 mantissa at a per-array power-of-two scale (5e-4 relative, far inside the 2e-2 budget)."""
    import numpy as np
    a = np.asarray(a, dtype=np.float32)
    m = float(np.abs(a).max())
    e = 0 if m == 0 else int(np.ceil(np.log2(m / 32768.0)))
    return (a / np.float32(2.0 ** e)).astype(np.float16), np.int32(e)


def fs_unpack(h, e):
    import numpy as np
    return h.astype(np.float32) * np.float32(2.0 ** int(e))


# ---- §8f row 3: the fork's SRImageModel_Feat (unet_other.py:43-77) and the denoise_start_point / q_sample(img2) start
# of p_sample_loop (gaussian_diffusion.py:517-523); oracle/make_golden_fork.py -> tests/golden/fork_golden.npz ------
SRFEAT_SEED, SRFEAT_LOOP_SEED = 71, 72
SRFEAT_KW = dict(SR_KW)                       # sr_create_model(..., conditioning="clip_feat")
SRFEAT_DIFFUSION = dict(_TR, timestep_respacing="250")
SRFEAT_START = 40                             # denoise_start_point: q_sample(img2, t=40) (SDEdit-style partial noising,
SRFEAT_RECORD = (1, 10, 40)                   # original t = 160), then steps 39 .. 0; samples recorded after these counts


def ref_srfeat_kwargs():
    kw = ref_sr_kwargs()
    kw.update(num_classes=512)
    return kw


def srfeat_inputs():
    g = th.Generator().manual_seed(INPUT_SEED + 4)
    x = th.randn(2, 3, IMAGE, IMAGE, generator=g)
    img2 = th.rand(2, 3, IMAGE, IMAGE, generator=g) * 2 - 1
    f1 = th.randn(2, 1, 512, generator=g)
    f2 = th.randn(2, 1, 512, generator=g)
    return (x, th.tensor([640, 15]), f1 / f1.norm(dim=-1, keepdim=True), f2 / f2.norm(dim=-1, keepdim=True), img2)
