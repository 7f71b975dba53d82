"""Diagnostic (not a test): UNet forward vs the CPU oracle at image sizes off the power-of-two grid, under the engine's
debug toggles, to localise a size-dependent fault.  python tests/diag_sizes.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th  # noqa: E402

from guided_diffusion_clip_b200 import _lib as L  # noqa: E402
from guided_diffusion_clip_b200 import script_util as su  # noqa: E402
from oracle import golden_cfg as cfg  # noqa: E402
from oracle import oracle_models as om  # noqa: E402


def run(size, attn, mult, env=None, dbg=None):
    for k, v in (env or {}).items():
        os.environ[k] = v
    lib = L.load()
    for k, v in (dbg or {}).items():
        lib.gd_debug_set(k, v)
    kw = dict(cfg.UNET_KW, image_size=size, attention_resolutions=attn, channel_mult=mult)
    m = su.create_model(**kw)
    sd = om.make_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 77)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    g = th.Generator().manual_seed(78)
    x = th.randn(3, 3, size, size, generator=g)
    t, y = th.tensor([5, 500, 999]), th.tensor([0, 10, 999])
    with th.no_grad():
        ref = om.unet_forward(sd, x, t, y, num_res_blocks=1, channel_mult_len=len(mult.split(",")), head_dim=64,
                              new_order=True)
        out = m(x.cuda(), t.cuda(), y.cuda()).cpu()
    err = float((out - ref).abs().max() / ref.abs().max())
    for k in (env or {}):
        os.environ.pop(k)
    for k in (dbg or {}):
        lib.gd_debug_set(k, 1 if k in (3, 4) else 0)
    return err


def run_clf(size, attn):
    kw = dict(cfg.CLASSIFIER_KW, image_size=size, classifier_attention_resolutions=attn)
    m = su.create_classifier(**kw)
    sd = om.make_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 79)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    g = th.Generator().manual_seed(80)
    x = th.randn(3, 3, size, size, generator=g)
    t, y = th.tensor([5, 500, 999]), th.tensor([0, 10, 999])
    from guided_diffusion_clip_b200.sampler import ClassifierGuidance
    try:
        with th.no_grad():
            logits = m(x.cuda(), t.cuda()).cpu()
        grad = ClassifierGuidance(m, 1.0)(x.cuda(), t.cuda(), y=y.cuda()).cpu()
    except Exception as e:  # noqa: BLE001
        return f"raised {type(e).__name__}: {str(e)[:120]}"
    with th.no_grad():
        ref_l = om.classifier_forward(sd, x, t, **cfg.CLF_STRUCT)
    ref_g = om.classifier_guidance(sd, x, t, y, 1.0, **cfg.CLF_STRUCT)
    return (f"logits err {float((logits - ref_l).abs().max() / ref_l.abs().max()):.3e}, "
            f"grad err {float((grad - ref_g).abs().max() / ref_g.abs().max()):.3e}")


if __name__ == "__main__":
    for size, attn in [(64, "16,8")]:  # create_classifier accepts 64 / 128 / 256 / 512 only (script_util.py:244-255)
        print(f"classifier size {size} attn {attn}: {run_clf(size, attn)}", flush=True)
    cases = [(64, "16", "1,2,4"), (48, "12", "1,2,4"), (96, "24", "1,2,4"), (80, "20", "1,2,4"), (48, "12", "1,2"),
             (32, "8", "1,2,4"), (128, "32", "1,2,4")]
    for size, attn, mult in cases:
        print(f"size {size} attn {attn} mult {mult}: err {run(size, attn, mult):.3e}", flush=True)
    for name, env, dbg in [("no conv_in", {"GD_B200_NO_CONV_IN": "1"}, None),
                           ("no fused stats", {"GD_B200_NO_FUSED_STATS": "1"}, None),
                           ("no tma epilogue, no fused stats", {"GD_B200_NO_FUSED_STATS": "1"}, {2: 1}), ("halo off", None, {4: 0}), ("pair mode off", None, {3: 0})]:
        print(f"size 48 [{name}]: err {run(48, '12', '1,2,4', env, dbg):.3e}", flush=True)
