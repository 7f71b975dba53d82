"""ORACLE — TEST INFRASTRUCTURE ONLY.  Generates tests/golden/* by running the REAL reference
(/root/reference/guided_diffusion, imported read-only, never copied) in the build container:

    python oracle/make_golden.py

The reference cannot travel to the GPU box, so its outputs are committed as small fixtures:
  diffusion_golden.json : space_timesteps sets, ValueError cases, timestep maps, float64 tables
  models_golden.npz     : tiny UNet / classifier outputs, guidance gradient, per-step p_sample / ddim_sample
                          outputs and short guided trajectories, all on oracle.make_state_dict weights
Inputs are regenerated from seeds by the tests (torch CPU generators are deterministic for a fixed version).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch as th
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from guided_diffusion import gaussian_diffusion as rgd  # noqa: E402  (the reference)
from guided_diffusion import respace as rrs  # noqa: E402
from guided_diffusion import script_util as rsu  # noqa: E402
from guided_diffusion import unet as runet  # noqa: E402

from oracle import golden_cfg as cfg  # noqa: E402
from oracle.oracle_models import make_state_dict  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def diffusion_golden():
    g = {"space_timesteps": {}, "errors": {}, "maps": {}, "tables": {}}
    for T, spec in cfg.SPACE_CASES:
        g["space_timesteps"][f"{T}|{spec}"] = sorted(rrs.space_timesteps(T, spec))
    for T, spec in cfg.SPACE_ERRORS:
        try:
            rrs.space_timesteps(T, spec)
            g["errors"][f"{T}|{spec}"] = None
        except ValueError as e:
            g["errors"][f"{T}|{spec}"] = str(e)
    for name, kw in cfg.DIFFUSION_CASES.items():
        d = rsu.create_gaussian_diffusion(**kw)
        g["maps"][name] = list(d.timestep_map)
        g["tables"][name] = {
            "betas": d.betas.tolist(),
            "alphas_cumprod": d.alphas_cumprod.tolist(),
            "alphas_cumprod_prev": d.alphas_cumprod_prev.tolist(),
            "sqrt_recip_alphas_cumprod": d.sqrt_recip_alphas_cumprod.tolist(),
            "sqrt_recipm1_alphas_cumprod": d.sqrt_recipm1_alphas_cumprod.tolist(),
            "posterior_variance": d.posterior_variance.tolist(),
            "posterior_log_variance_clipped": d.posterior_log_variance_clipped.tolist(),
            "posterior_mean_coef1": d.posterior_mean_coef1.tolist(),
            "posterior_mean_coef2": d.posterior_mean_coef2.tolist(),
            "model_mean_type": d.model_mean_type.name, "model_var_type": d.model_var_type.name,
            "loss_type": d.loss_type.name, "rescale_timesteps": bool(d.rescale_timesteps),
            "num_timesteps": int(d.num_timesteps),
        }
    g["defaults"] = {
        "diffusion_defaults": rsu.diffusion_defaults(),
        "classifier_defaults": rsu.classifier_defaults(),
        "model_and_diffusion_defaults": rsu.model_and_diffusion_defaults(),
        "sr_model_and_diffusion_defaults": rsu.sr_model_and_diffusion_defaults(),
    }
    import hashlib

    def layout(sd):
        h = hashlib.sha256()
        for k, v in sd.items():
            h.update(f"{k}:{tuple(v.shape)};".encode())
        return [h.hexdigest()[:16], len(sd)]

    with th.device("meta"):
        u256 = runet.UNetModel(**cfg.ref_unet256_kwargs())
        c256 = rsu.create_classifier(**cfg.CLF256_KW)
        sr = runet.SuperResModel(**cfg.ref_sr512_kwargs())
        ut = runet.UNetModel(**cfg.ref_unet_kwargs())
        ct = rsu.create_classifier(**cfg.CLASSIFIER_KW)
    g["layouts"] = {
        "unet_tiny": layout(ut.state_dict()), "clf_tiny": layout(ct.state_dict()),
        "unet_256": layout(u256.state_dict()), "clf_256": layout(c256.state_dict()), "sr_512": layout(sr.state_dict()),
        "unet_256_params": int(sum(p.numel() for p in u256.parameters())),
    }
    with open(os.path.join(OUT, "diffusion_golden.json"), "w") as f:
        json.dump(g, f)


def build_ref_unet():
    m = runet.UNetModel(**cfg.ref_unet_kwargs())
    sd = make_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, cfg.UNET_SEED)
    m.load_state_dict(sd, strict=True)
    return m.eval(), sd


def build_ref_classifier():
    m = rsu.create_classifier(**cfg.CLASSIFIER_KW)
    sd = make_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, cfg.CLF_SEED)
    m.load_state_dict(sd, strict=True)
    return m.eval(), sd


def models_golden():
    out = {}
    th.set_grad_enabled(False)
    unet, _ = build_ref_unet()
    clf, _ = build_ref_classifier()
    x, t, y = cfg.model_inputs()
    out["unet_out"] = unet(x, t, y).numpy()
    out["clf_logits"] = clf(x, t).numpy()

    def cond_fn(x, t, y=None):
        with th.enable_grad():
            x_in = x.detach().requires_grad_(True)
            logits = clf(x_in, t)
            log_probs = F.log_softmax(logits, dim=-1)
            selected = log_probs[range(len(logits)), y.view(-1)]
            return th.autograd.grad(selected.sum(), x_in)[0] * cfg.CLF_SCALE

    out["clf_grad"] = cond_fn(x, t, y=y).numpy()

    # §8f rows: SuperResModel (bilinear low_res conditioning) and the fork's UNetModel_clip_feat
    from guided_diffusion import unet_other as rother
    sr = runet.SuperResModel(**cfg.ref_sr_kwargs())
    sr.load_state_dict(make_state_dict({k: tuple(v.shape) for k, v in sr.state_dict().items()}, cfg.SR_SEED), strict=True)
    xs_, ts_, ys_, low_ = cfg.sr_inputs()
    out["sr_out"] = sr.eval()(xs_, ts_, low_res=low_, y=ys_).numpy()
    fm = rother.UNetModel_clip_feat(**cfg.ref_feat_kwargs())
    fm.load_state_dict(make_state_dict({k: tuple(v.shape) for k, v in fm.state_dict().items()}, cfg.FEAT_SEED), strict=True)
    xf_, tf_, feat_ = cfg.feat_inputs()
    out["feat_out"] = fm.eval()(xf_, tf_, clip_feat=feat_).numpy()

    # per-step update with a fixed fake model output (pure diffusion arithmetic)
    for name, kw in cfg.STEP_CASES.items():
        d = rsu.create_gaussian_diffusion(**kw["diffusion"])
        xs, mo, g, i = cfg.step_inputs(name)
        tt = th.tensor([i] * xs.shape[0])
        fake_model = lambda x_, t_, **k: mo
        fake_cond = (lambda x_, t_, **k: g) if kw["guided"] else None
        th.manual_seed(cfg.STEP_NOISE_SEED)
        if kw["ddim"]:
            r = d.ddim_sample(fake_model, xs, tt, cond_fn=fake_cond, model_kwargs={}, eta=kw["eta"])
        else:
            r = d.p_sample(fake_model, xs, tt, cond_fn=fake_cond, model_kwargs={})
        pmv = d.p_mean_variance(fake_model, xs, tt, model_kwargs={})
        out[f"step_{name}_sample"] = r["sample"].numpy()
        out[f"step_{name}_x0"] = r["pred_xstart"].numpy()
        out[f"step_{name}_mean"] = pmv["mean"].numpy()
        out[f"step_{name}_var"] = pmv["variance"].numpy()
        out[f"step_{name}_logvar"] = pmv["log_variance"].numpy()

    # short guided trajectories through the real reference loops (CPU generator noise)
    for name, kw in cfg.TRAJ_CASES.items():
        d = rsu.create_gaussian_diffusion(**kw["diffusion"])
        ts_seen = []

        def model_fn(x_, t_, y=None):
            ts_seen.append(int(t_[0]))
            return unet(x_, t_, y)

        th.manual_seed(cfg.TRAJ_SEED)
        yy = cfg.traj_labels()
        fn = d.ddim_sample_loop if kw["ddim"] else d.p_sample_loop
        s = fn(model_fn, (cfg.TRAJ_BATCH, 3, cfg.IMAGE, cfg.IMAGE), model_kwargs={"y": yy},
               cond_fn=cond_fn if kw["guided"] else None, device="cpu")
        out[f"traj_{name}"] = s.numpy()
        out[f"traj_{name}_ts"] = np.array(ts_seen, dtype=np.int64)
        out[f"traj_{name}_u8"] = ((s + 1) * 127.5).clamp(0, 255).to(th.uint8).permute(0, 2, 3, 1).contiguous().numpy()
    # first 10 steps of the full-length chains through the reference's progressive generators
    for name, kw in cfg.TRAJ10_CASES.items():
        d = rsu.create_gaussian_diffusion(**kw["diffusion"])
        th.manual_seed(cfg.TRAJ_SEED)
        yy = cfg.traj_labels()
        gen = (d.ddim_sample_loop_progressive if kw["ddim"] else d.p_sample_loop_progressive)(
            lambda x_, t_, y=None: unet(x_, t_, y), (cfg.TRAJ_BATCH, 3, cfg.IMAGE, cfg.IMAGE), model_kwargs={"y": yy},
            cond_fn=cond_fn if kw["guided"] else None, device="cpu", **({} if kw["ddim"] else {"denoise_start_point": -1}))
        last = None
        for k, o in enumerate(gen):
            last = o
            if k + 1 == cfg.TRAJ10_STEPS:
                break
        out[f"traj10_{name}_sample"] = last["sample"].numpy()
        out[f"traj10_{name}_x0"] = last["pred_xstart"].numpy()
    np.savez_compressed(os.path.join(OUT, "models_golden.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, float(np.abs(v).max()) if v.dtype != np.uint8 else "")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    diffusion_golden()
    models_golden()
    print("golden fixtures written to", OUT)
