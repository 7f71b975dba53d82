"""ORACLE — TEST INFRASTRUCTURE ONLY.  Full-size one-step fixtures of BASELINE configs[1..4], produced by the REAL
reference (/root/reference/guided_diffusion imported read-only; the CLIP encoder of configs[2] by
transformers.CLIPVisionModelWithProjection, the stand-in reference of SURVEY §8c) in fp32 on the build container's CPU:

    python -m oracle.make_golden_fullsize [cfg2 cfg3 cfg4 cfg5]      ->  tests/golden/fullsize_<case>.npz

Each case: batch 1, oracle.make_state_dict weights (seeded, nothing zero), one guided step through the reference's own
p_sample / ddim_sample (scripts/classifier_sample.py:54-65 closures), recording the model output eps|v
(unet.py:635-664), the guidance gradient (classifier_sample.py:54-61 / the CLIP cosine gradient) and x_{t-1}, pred_xstart
(gaussian_diffusion.py:395-439, 546-594).  The GPU tests rebuild weights and inputs from the same seeds."""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch as th
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from guided_diffusion import script_util as rsu  # noqa: E402  (the reference)
from guided_diffusion import unet as runet  # noqa: E402

from oracle import golden_cfg as cfg  # noqa: E402
from oracle import oracle_clip as oc  # noqa: E402
from oracle.oracle_models import make_state_dict  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _load(m, seed):
    sd = make_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed)
    m.load_state_dict(sd, strict=True)
    return m.eval()


def build_models(name):
    c = cfg.FULLSIZE_CASES[name]
    seed = cfg.FS_SEED + c["k"]
    clf = clip = None
    if name == "cfg2":
        unet = runet.UNetModel(**dict(cfg.ref_unet256_kwargs(), use_fp16=False))
        clf = rsu.create_classifier(**cfg.CLF256_KW)
    elif name == "cfg3":
        unet = runet.UNetModel(**dict(cfg.ref_unet256_kwargs(), use_fp16=False, num_classes=None))
        from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
        clip = CLIPVisionModelWithProjection(CLIPVisionConfig(hidden_act="quick_gelu", **cfg.CLIP_B16)).eval()
        clip.load_state_dict(cfg.clip_state_dict({k: tuple(v.shape) for k, v in clip.state_dict().items()}), strict=True)
    elif name == "cfg4":
        unet = runet.SuperResModel(**dict(cfg.ref_sr512_kwargs(), use_fp16=False))
    else:
        unet = runet.UNetModel(**cfg.ref_unet512_kwargs())
        clf = rsu.create_classifier(**cfg.CLF512_KW)
    _load(unet, seed)
    if clf is not None:
        _load(clf, seed + 100)
    return unet, clf, clip


def run_case(name):
    c = cfg.FULLSIZE_CASES[name]
    t0 = time.time()
    unet, clf, clip = build_models(name)
    x, low, txt = cfg.fullsize_inputs(name)
    d = rsu.create_gaussian_diffusion(**c["diffusion"])
    tt = th.tensor([c["index"]])
    y = th.tensor([c["label"]]) if c["label"] is not None else None
    rec = {}

    def model_fn(x_, t_, **kw):
        with th.no_grad():
            if name == "cfg4":
                o = unet(x_, t_, low_res=kw["low_res"], y=kw["y"])
            elif name == "cfg3":
                o = unet(x_, t_)
            else:
                o = unet(x_, t_, kw["y"])
        rec["eps"] = o.detach().clone()
        rec["t_model"] = int(t_[0])
        return o

    def cond_clf(x_, t_, y=None, **kw):  # scripts/classifier_sample.py:54-61
        with th.enable_grad():
            x_in = x_.detach().requires_grad_(True)
            logits = clf(x_in, t_)
            log_probs = F.log_softmax(logits, dim=-1)
            selected = log_probs[range(len(logits)), y.view(-1)]
            g = th.autograd.grad(selected.sum(), x_in)[0] * c["scale"]
        rec["grad"], rec["logits"] = g.detach().clone(), logits.detach().clone()
        return g

    def cond_clip(x_, t_, **kw):  # SURVEY §8c spec
        with th.enable_grad():
            x_in = x_.detach().requires_grad_(True)
            e = clip(pixel_values=oc.preprocess(x_in, cfg.CLIP_B16["image_size"])).image_embeds
            e = e / e.norm(dim=-1, keepdim=True)
            sim = c["scale"] * (e * txt).sum(-1)
            g = th.autograd.grad(sim.sum(), x_in)[0]
        rec["grad"], rec["sim"] = g.detach().clone(), sim.detach().clone()
        return g

    cond = cond_clf if clf is not None else cond_clip if clip is not None else None
    mk = {}
    if y is not None:
        mk["y"] = y
    if name == "cfg4":
        mk["low_res"] = low
    th.manual_seed(cfg.FS_SEED + 20 + c["k"])
    with th.no_grad():
        if c["ddim"]:
            r = d.ddim_sample(model_fn, x, tt, cond_fn=cond, model_kwargs=mk, eta=0.0)
        else:
            r = d.p_sample(model_fn, x, tt, cond_fn=cond, model_kwargs=mk)
    out = {"t_model": np.int64(rec["t_model"])}
    for key, val in (("eps", rec["eps"]), ("grad", rec.get("grad")), ("sample", r["sample"]), ("x0", r["pred_xstart"]),
                     ("logits", rec.get("logits")), ("sim", rec.get("sim"))):
        if val is None:
            continue
        if key in ("logits", "sim"):
            out[key] = val.numpy().astype(np.float32)
        else:
            out[key], out[key + "_exp"] = cfg.fs_pack(val.numpy())
    path = os.path.join(OUT, f"fullsize_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {time.time() - t0:.1f} s, {os.path.getsize(path) / 1e6:.2f} MB;",
          {k: (tuple(v.shape), float(np.abs(v.astype(np.float32)).max())) for k, v in out.items() if getattr(v, 'ndim', 0)},
          flush=True)


def run_traj():
    """north_star: "short 10-step trajectories must match within the same stated bound" — at the REAL widths: the first
    10 reverse steps of the 250-step classifier-guided chain of configs[1] through the reference's own
    p_sample_loop_progressive (batch 1, fp32, CPU-generator noise), recording x after 5 and 10 steps."""
    t0 = time.time()
    unet, clf, _ = build_models("cfg2")
    c = cfg.FULLSIZE_CASES["cfg2"]
    d = rsu.create_gaussian_diffusion(**c["diffusion"])
    y = th.tensor([c["label"]])

    def model_fn(x_, t_, y=None):
        return unet(x_, t_, y)

    def cond_fn(x_, t_, y=None):
        with th.enable_grad():
            x_in = x_.detach().requires_grad_(True)
            log_probs = F.log_softmax(clf(x_in, t_), dim=-1)
            return th.autograd.grad(log_probs[range(len(x_in)), y.view(-1)].sum(), x_in)[0] * c["scale"]

    th.manual_seed(cfg.FS_SEED + 30)
    out = {}
    gen = d.p_sample_loop_progressive(model_fn, (1, 3, 256, 256), model_kwargs={"y": y}, cond_fn=cond_fn, device="cpu",
                                      denoise_start_point=-1)
    for k, o in enumerate(gen):
        if k + 1 in (5, cfg.FS_TRAJ_STEPS):
            out[f"sample{k + 1}"], out[f"sample{k + 1}_exp"] = cfg.fs_pack(o["sample"].numpy())
        if k + 1 == cfg.FS_TRAJ_STEPS:
            break
    path = os.path.join(OUT, "fullsize_traj_cfg2.npz")
    np.savez_compressed(path, **out)
    print(f"traj cfg2: {time.time() - t0:.1f} s, {os.path.getsize(path) / 1e6:.2f} MB", flush=True)


if __name__ == "__main__":
    th.set_num_threads(os.cpu_count())
    names = sys.argv[1:] or (list(cfg.FULLSIZE_CASES) + ["traj"])
    for nm in names:
        if nm == "traj":
            run_traj()
        else:
            run_case(nm)
