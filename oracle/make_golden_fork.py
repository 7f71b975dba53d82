"""ORACLE — TEST INFRASTRUCTURE ONLY.  Golden vectors for the fork's own use-case, produced by the
REAL reference (/root/reference imported read-only):   python -m oracle.make_golden_fork  -> tests/golden/fork_golden.npz

  srfeat_out    SRImageModel_Feat.forward (unet_other.py:43-77): x ‖ img2 input, y = clip_feat − clip_feat2 + bias_feat
  srfeat_loop   p_sample_loop(..., denoise_start_point=40 of 250) (gaussian_diffusion.py:517-523): the chain starts from
                q_sample(model_kwargs['img2'], t=40) instead of noise; CPU-generator draws in the reference's order
                (randn(shape), q_sample's randn_like, one randn_like per step)."""
import os
import sys

import numpy as np
import torch as th

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from guided_diffusion import script_util as rsu  # noqa: E402  (the reference)
from guided_diffusion import unet_other as rother  # noqa: E402

from oracle import golden_cfg as cfg  # noqa: E402
from oracle.oracle_models import make_state_dict  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    th.set_grad_enabled(False)
    m = rother.SRImageModel_Feat(**cfg.ref_srfeat_kwargs())
    m.load_state_dict(make_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, cfg.SRFEAT_SEED),
                      strict=True)
    m.eval()
    x, t, f1, f2, img2 = cfg.srfeat_inputs()
    out = {"srfeat_out": m(x, t, clip_feat=f1, clip_feat2=f2, img2=img2).numpy()}
    d = rsu.create_gaussian_diffusion(**cfg.SRFEAT_DIFFUSION)
    ts = []

    def model_fn(x_, t_, **kw):
        ts.append(int(t_[0]))
        return m(x_, t_, **kw)

    th.manual_seed(cfg.SRFEAT_LOOP_SEED)
    steps = [o["sample"].numpy() for o in d.p_sample_loop_progressive(
        model_fn, tuple(x.shape), model_kwargs={"clip_feat": f1, "clip_feat2": f2, "img2": img2}, device="cpu",
        denoise_start_point=cfg.SRFEAT_START)]
    out["srfeat_loop"] = np.stack([steps[k - 1] for k in cfg.SRFEAT_RECORD])  # the last one is what p_sample_loop returns
    print("saturated fraction of the final sample:", float((np.abs(steps[-1]) >= 1.0).mean()))
    out["srfeat_loop_ts"] = np.array(ts, dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "fork_golden.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, v if v.dtype == np.int64 else float(np.abs(v).max()))


if __name__ == "__main__":
    main()
