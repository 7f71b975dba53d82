"""ORACLE — TEST INFRASTRUCTURE ONLY.  BASELINE configs[0] at full size through the REAL reference
(/root/reference, imported read-only): 64x64 class-conditional ADM, unguided p_sample_loop_progressive,
timestep_respacing "25", batch 4, on CPU.

    python oracle/make_golden_config1.py   ->  tests/golden/config1_golden.npz  (+ prints the CPU wall time)
"""
from __future__ import annotations

import os
import sys
import time

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
    m = runet.UNetModel(**cfg.ref_c1_kwargs())
    sd = make_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, cfg.C1_SEED)
    m.load_state_dict(sd, strict=True)
    m.eval()
    d = rsu.create_gaussian_diffusion(**cfg.C1_DIFFUSION)
    y = cfg.c1_labels()
    th.manual_seed(cfg.C1_NOISE_SEED)  # the loop draws randn(shape) and one randn_like per step: cfg.c1_noise()
    out = {}
    t0 = time.time()
    gen = d.p_sample_loop_progressive(lambda x, t, y=None: m(x, t, y), (cfg.C1_BATCH, 3, 64, 64),
                                      model_kwargs={"y": y}, device="cpu", denoise_start_point=-1)
    for k, o in enumerate(gen, 1):
        if k in cfg.C1_CHECKPOINTS:
            out[f"sample_after_{k}"] = o["sample"].numpy()
            out[f"x0_after_{k}"] = o["pred_xstart"].numpy()
    dt = time.time() - t0
    assert k == cfg.C1_STEPS
    out["cpu_seconds"] = np.array([dt])
    out["cpu_threads"] = np.array([th.get_num_threads()])
    path = os.path.join(ROOT, "tests", "golden", "config1_golden.npz")
    np.savez_compressed(path, **out)
    for kk, v in out.items():
        print(kk, v.shape, float(np.abs(v).max()))
    print(f"reference config 1 on CPU: {dt:.1f} s for {cfg.C1_BATCH} samples, {th.get_num_threads()} threads ->", path,
          os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
