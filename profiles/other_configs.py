"""Step time of BASELINE configs[0] and configs[2..4] through the public sampling API (one process, one B200),
CUDA-event timed:

  cfg1  64x64 class-cond ADM (192 ch, 3 res blocks), unguided p_sample_loop, respacing 25, batch 4 (and 256): the
        WHOLE loop through diffusion.p_sample_loop, wall clock, next to the reference's CPU time for the same job

  cfg3  256x256 unconditional ADM + CLIP ViT-B/16 image-encoder guidance, DDIM-50, batch 32
  cfg4  128->512 upsampler (SuperResModel, 192 ch), 250 steps, batch 8
  cfg5  512x512 class-cond ADM (use_fp16=False masters) + classifier-512 guidance (scale 4.0), DDIM-25, batch 8/16

These are parity-test configurations, not bench lines (bench.py measures configs[1]); the numbers go into DESIGN.md.
TFLOP/s uses SURVEY 8(d)'s algorithmic FLOPs per sample and model evaluation.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch as th  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import bench  # noqa: E402
from guided_diffusion_clip_b200 import clip as gclip  # noqa: E402
from guided_diffusion_clip_b200 import script_util as su  # noqa: E402
from guided_diffusion_clip_b200.sampler import ClassifierGuidance, ModelFn  # noqa: E402

dev = th.device("cuda", 0)


def time_steps(fn, steps=6, warm=2):
    for _ in range(warm):
        fn()
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    th.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def cfg1(batch=4):
    import time
    kw = su.model_and_diffusion_defaults()
    kw.update(image_size=64, num_channels=192, num_res_blocks=3, attention_resolutions="32,16,8", num_head_channels=64,
              class_cond=True, learn_sigma=True, noise_schedule="cosine", resblock_updown=True,
              use_scale_shift_norm=True, use_new_attention_order=True, timestep_respacing="25")
    model, diffusion = su.create_model_and_diffusion(**kw)
    bench.randomize_(model, 3)
    model.to(dev).eval()
    g = th.Generator(device="cuda").manual_seed(4)
    y = th.randint(0, 1000, (batch,), generator=g, device=dev)
    mf = ModelFn(model, True)

    def loop():
        with th.no_grad():
            return diffusion.p_sample_loop(mf, (batch, 3, 64, 64), model_kwargs={"y": y})

    loop()  # plan build + graph capture
    th.cuda.synchronize()
    t0 = time.time()
    reps = 3
    for _ in range(reps):
        out = loop()
    th.cuda.synchronize()
    s = (time.time() - t0) / reps
    assert bool(th.isfinite(out).all())
    return {"config": "cfg1 class-cond-64 (192ch x3), unguided p_sample_loop, 25 steps", "batch": batch,
            "loop_seconds": round(s, 4), "ms_per_step": round(1e3 * s / 25, 3), "samples_per_s": round(batch / s, 2),
            "step_tflops": round(219.36 * batch / (1e3 * s / 25), 1),
            "reference_cpu_seconds_batch4": "59.8 s on 8 threads in the build container (oracle/make_golden_config1.py)"}


def cfg3(batch=32):
    kw = bench.unet_kwargs(256)
    kw.update(class_cond=False, timestep_respacing="ddim50")
    model, diffusion = su.create_model_and_diffusion(**kw)
    bench.randomize_(model, 11)
    model.to(dev).convert_to_fp16()
    model.eval()
    enc = gclip.CLIPVisionEncoder().to(dev).eval()
    g = th.Generator(device="cuda").manual_seed(5)
    txt = F.normalize(th.randn((1, 512), generator=g, device=dev), dim=-1)
    cond = gclip.CLIPGuidance(enc, txt, 100.0)
    x = th.randn((batch, 3, 256, 256), generator=g, device=dev)
    t = th.full((batch,), 30, device=dev, dtype=th.int64)

    def step():
        with th.no_grad():
            return diffusion.ddim_sample(model, x, t, cond_fn=cond, model_kwargs={})

    ms = time_steps(step)
    ms_clip = time_steps(lambda: cond(x, None))
    gf = 2239.67 + 3 * 35.1
    return {"config": "cfg3 uncond-256 + CLIP ViT-B/16 guidance, DDIM-50", "batch": batch, "ms_per_step": round(ms, 2),
            "clip_fwd_bwd_ms": round(ms_clip, 2), "samples_per_s": round(batch / (50 * ms / 1e3), 3),
            "step_tflops": round(gf * batch / ms, 1)}


def cfg4(batch=8):
    kw = su.sr_model_and_diffusion_defaults()
    kw.update(large_size=512, small_size=128, num_channels=192, num_res_blocks=2, attention_resolutions="32,16",
              num_head_channels=64, class_cond=True, learn_sigma=True, resblock_updown=True, use_scale_shift_norm=True,
              use_fp16=True, timestep_respacing="250")
    model, diffusion = su.sr_create_model_and_diffusion(**kw)
    bench.randomize_(model, 99)
    model.to(dev).convert_to_fp16()
    model.eval()
    g = th.Generator(device="cuda").manual_seed(6)
    x = th.randn((batch, 3, 512, 512), generator=g, device=dev)
    low = th.rand((batch, 3, 128, 128), generator=g, device=dev) * 2 - 1
    t = th.full((batch,), 100, device=dev, dtype=th.int64)
    y = th.randint(0, 1000, (batch,), generator=g, device=dev)

    def step():
        with th.no_grad():
            return diffusion.p_sample(model, x, t, model_kwargs={"low_res": low, "y": y})

    ms = time_steps(step)
    return {"config": "cfg4 128->512 upsampler, 250 steps", "batch": batch, "ms_per_step": round(ms, 2),
            "samples_per_s": round(batch / (250 * ms / 1e3), 4), "step_tflops": round(5009.57 * batch / ms, 1)}


def cfg5(batch=8):
    kw = bench.unet_kwargs(512)
    kw.update(use_fp16=False, timestep_respacing="ddim25")
    model, diffusion = su.create_model_and_diffusion(**kw)
    bench.randomize_(model, 7)
    model.to(dev).eval()
    ckw = bench.clf_kwargs(512)
    ckw.update(classifier_use_fp16=False)
    clf = su.create_classifier(**ckw)
    bench.randomize_(clf, 8)
    clf.to(dev).eval()
    cond = ClassifierGuidance(clf, 4.0)
    mf = ModelFn(model, True)
    g = th.Generator(device="cuda").manual_seed(9)
    x = th.randn((batch, 3, 512, 512), generator=g, device=dev)
    t = th.full((batch,), 12, device=dev, dtype=th.int64)
    y = th.randint(0, 1000, (batch,), generator=g, device=dev)

    def step():
        with th.no_grad():
            return diffusion.ddim_sample(mf, x, t, cond_fn=cond, model_kwargs={"y": y})

    ms = time_steps(step)
    gf = 3964.67 + 2 * 225.53
    return {"config": "cfg5 class-cond-512 + classifier-512 guidance (scale 4), DDIM-25", "batch": batch,
            "ms_per_step": round(ms, 2), "samples_per_s": round(batch / (25 * ms / 1e3), 3),
            "step_tflops": round(gf * batch / ms, 1)}


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg1", "cfg1:256", "cfg3", "cfg4", "cfg5"]
    for name in which:
        fn, _, b = name.partition(":")
        r = globals()[fn](int(b)) if b else globals()[fn]()
        print(json.dumps(r), flush=True)
        th.cuda.empty_cache()
