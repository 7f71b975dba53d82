"""Golden vectors for the CLIP image-encoder guidance (tests/golden/clip_golden.npz).

Run in the BUILD container (needs `transformers`, the stand-in reference named in SURVEY §8c — the reference
repository itself has no CLIP code):   python -m oracle.make_golden_clip
Pins oracle/oracle_clip.py against transformers.CLIPVisionModelWithProjection on the same seeded weights."""
import os

import numpy as np
import torch as th

from oracle import golden_cfg as cfg
from oracle import oracle_clip as oc

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    th.manual_seed(0)
    hf = CLIPVisionModelWithProjection(CLIPVisionConfig(hidden_act="quick_gelu", **cfg.CLIP_TINY)).eval()
    sd = cfg.clip_state_dict({k: tuple(v.shape) for k, v in hf.state_dict().items()})
    hf.load_state_dict(sd, strict=True)
    x, txt = cfg.clip_inputs()
    out = {}
    pix = oc.preprocess(x, cfg.CLIP_TINY["image_size"])
    with th.no_grad():
        out["clip_pixels"] = pix.numpy()
        out["clip_embed"] = hf(pixel_values=pix).image_embeds.numpy()
    xin = x.clone().requires_grad_(True)
    e = hf(pixel_values=oc.preprocess(xin, cfg.CLIP_TINY["image_size"])).image_embeds
    e = e / e.norm(dim=-1, keepdim=True)
    sim = cfg.CLIP_SCALE * (e * txt).sum(-1)
    out["clip_sim"] = sim.detach().numpy()
    out["clip_grad"] = th.autograd.grad(sim.sum(), xin)[0].numpy()
    np.savez_compressed(os.path.join(OUT, "clip_golden.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, float(np.abs(v).max()))


if __name__ == "__main__":
    main()
