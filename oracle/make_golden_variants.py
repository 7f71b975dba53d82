"""ORACLE — TEST INFRASTRUCTURE ONLY.  Golden outputs of the REAL reference (/root/reference, imported read-only)
for the model variants of golden_cfg.VARIANT_KW (factory-default heads / conv resampling / additive embedding) and
for the denoised_fn hook of p_sample / ddim_sample:

    python oracle/make_golden_variants.py   ->  tests/golden/variants_golden.npz
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch as th

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from guided_diffusion import script_util as rsu  # noqa: E402  (the reference)
from guided_diffusion import unet as runet  # noqa: E402

from oracle import golden_cfg as cfg  # noqa: E402
from oracle.oracle_models import make_state_dict  # noqa: E402


def main():
    th.set_grad_enabled(False)
    out = {}
    for i, name in enumerate(sorted(cfg.VARIANT_KW)):
        m = runet.UNetModel(**cfg.ref_variant_kwargs(name))
        sd = make_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, cfg.VAR_SEED + i)
        m.load_state_dict(sd, strict=True)
        x, t, y = cfg.variant_inputs(name)
        out[f"variant_{name}_out"] = m.eval()(x, t, y).numpy()
        out[f"variant_{name}_keys"] = np.array(list(sd.keys()))
    for name in cfg.DENOISED_CASES:
        kw = cfg.STEP_CASES[name]
        d = rsu.create_gaussian_diffusion(**kw["diffusion"])
        xs, mo, g, i = cfg.step_inputs(name)
        tt = th.tensor([i] * xs.shape[0])
        fake_model = lambda x_, t_, **k: mo  # noqa: E731
        fake_cond = (lambda x_, t_, **k: g) if kw["guided"] else None  # noqa: E731
        th.manual_seed(cfg.STEP_NOISE_SEED)
        if kw["ddim"]:
            r = d.ddim_sample(fake_model, xs, tt, denoised_fn=cfg.denoised_fn_example, cond_fn=fake_cond,
                              model_kwargs={}, eta=kw["eta"])
        else:
            r = d.p_sample(fake_model, xs, tt, denoised_fn=cfg.denoised_fn_example, cond_fn=fake_cond, model_kwargs={})
        pmv = d.p_mean_variance(fake_model, xs, tt, denoised_fn=cfg.denoised_fn_example, model_kwargs={})
        out[f"denoised_{name}_sample"] = r["sample"].numpy()
        out[f"denoised_{name}_x0"] = r["pred_xstart"].numpy()
        out[f"denoised_{name}_mean"] = pmv["mean"].numpy()
    for name, (step, i) in cfg.REVERSE_CASES.items():
        d = rsu.create_gaussian_diffusion(**cfg.STEP_CASES[step]["diffusion"])
        xs, mo, _, _ = cfg.step_inputs(step)
        r = d.ddim_reverse_sample(lambda x_, t_, **k: mo, xs, th.tensor([i] * xs.shape[0]), model_kwargs={})
        out[f"reverse_{name}_sample"] = r["sample"].numpy()
        out[f"reverse_{name}_x0"] = r["pred_xstart"].numpy()
    # classifier variant: logits and the guidance gradient of scripts/classifier_sample.py:54-61
    import torch.nn.functional as F
    for tag, kw, seed in (("plain", cfg.CLF_PLAIN_KW, cfg.CLF_PLAIN_SEED),
                          ("convdown", cfg.CLF_CONVDOWN_KW, cfg.CLF_CONVDOWN_SEED)):
        clf = rsu.create_classifier(**kw)
        clf.load_state_dict(make_state_dict({k: tuple(v.shape) for k, v in clf.state_dict().items()}, seed), strict=True)
        clf.eval()
        x, t, y = cfg.model_inputs()
        out[f"clf_{tag}_logits"] = clf(x, t).numpy()
        with th.enable_grad():
            x_in = x.detach().requires_grad_(True)
            sel = F.log_softmax(clf(x_in, t), dim=-1)[range(len(x)), y.view(-1)]
            out[f"clf_{tag}_grad"] = th.autograd.grad(sel.sum(), x_in)[0].numpy()
    path = os.path.join(ROOT, "tests", "golden", "variants_golden.npz")
    np.savez_compressed(path, **out)
    for k, v in out.items():
        print(k, v.shape, float(np.abs(v).max()) if v.dtype.kind == "f" else "")
    print("written", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
