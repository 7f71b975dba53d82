"""The generic CUDA-graph stepper (sampler.GenericGraphedStepper): every combination of OUR model classes and guidance
objects that the specialised classifier-guided stepper does not cover takes one graph replay per step — SuperResModel
with `low_res`, the fork's clip_feat conditioning, CLIP image-encoder guidance — and produces the same BITS as the
eager launch sequence (same kernels, same order, same RNG draws)."""
import pytest
import torch as th

from guided_diffusion_clip_b200 import clip as gclip
from guided_diffusion_clip_b200 import script_util as su
from guided_diffusion_clip_b200.sampler import GenericGraphedStepper, GraphedStepper, ModelFn
from oracle import golden_cfg as cfg
from oracle import oracle_models as om

pytestmark = pytest.mark.gpu


def _load(m, seed):
    m.load_state_dict(om.make_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed), strict=True)
    return m.cuda().eval()


def _both(monkeypatch, run):
    outs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("GD_B200_NO_GRAPH", flag)
        th.manual_seed(5)
        outs.append(run())
    return outs


def test_superres_loop_graph_equals_eager(lib, monkeypatch):
    m = _load(su.sr_create_model(**cfg.SR_KW), cfg.SR_SEED)
    d = su.create_gaussian_diffusion(steps=1000, learn_sigma=True, noise_schedule="linear", timestep_respacing="4")
    x, _, y, low = (v.cuda() for v in cfg.sr_inputs())
    kw = {"low_res": low, "y": y}
    a, b = _both(monkeypatch, lambda: d.p_sample_loop(m, tuple(x.shape), model_kwargs=kw, device="cuda"))
    assert th.equal(a, b) and bool(th.isfinite(a).all())
    st = GraphedStepper.cached(d, m, None, tuple(x.shape), "cuda", kw, True, False, 0.0)
    assert isinstance(st, GenericGraphedStepper) and st.launches_per_step > 100
    # new conditioning tensors of the same shape reuse the captured graph and are honoured
    kw2 = {"low_res": low.flip(0).contiguous(), "y": y}
    assert GraphedStepper.cached(d, m, None, tuple(x.shape), "cuda", kw2, True, False, 0.0) is st
    c, e = _both(monkeypatch, lambda: d.p_sample_loop(m, tuple(x.shape), model_kwargs=kw2, device="cuda"))
    assert th.equal(c, e) and not th.equal(a, c)


def test_clip_feat_ddim_loop_graph_equals_eager(lib, monkeypatch):
    m = _load(su.create_model(**cfg.FEAT_KW, conditioning="clip_feat"), cfg.FEAT_SEED)
    d = su.create_gaussian_diffusion(steps=1000, learn_sigma=True, noise_schedule="linear", timestep_respacing="ddim4")
    x, _, feat = (v.cuda() for v in cfg.feat_inputs())
    kw = {"clip_feat": feat}
    a, b = _both(monkeypatch, lambda: d.ddim_sample_loop(m, tuple(x.shape), model_kwargs=kw, device="cuda", eta=0.5))
    assert th.equal(a, b) and bool(th.isfinite(a).all())


def test_clip_guided_loop_graph_equals_eager(lib, monkeypatch):
    enc = gclip.CLIPVisionEncoder(**cfg.CLIP_TINY)
    enc.load_state_dict(cfg.clip_state_dict({k: tuple(v.shape) for k, v in enc.state_dict().items()}), strict=True)
    enc = enc.cuda().eval()
    unet = _load(su.create_model(**dict(cfg.UNET_KW, class_cond=False)), cfg.UNET_SEED + 100)
    d = su.create_gaussian_diffusion(steps=1000, learn_sigma=True, noise_schedule="linear", timestep_respacing="ddim4")
    _, txt = cfg.clip_inputs()
    cond = gclip.CLIPGuidance(enc, txt.cuda(), 20.0 * cfg.CLIP_SCALE)
    shape = (2, 3, cfg.IMAGE, cfg.IMAGE)
    a, b = _both(monkeypatch, lambda: d.ddim_sample_loop(unet, shape, cond_fn=cond, model_kwargs={}, device="cuda"))
    assert th.equal(a, b) and bool(th.isfinite(a).all())
    assert isinstance(GraphedStepper.cached(d, unet, cond, shape, "cuda", {}, True, True, 0.0), GenericGraphedStepper)


def test_closures_and_foreign_kwargs_stay_on_the_eager_path(lib):
    unet = _load(su.create_model(**cfg.UNET_KW), cfg.UNET_SEED)
    d = su.create_gaussian_diffusion(steps=1000, learn_sigma=True, noise_schedule="linear", timestep_respacing="4")
    shape = (2, 3, cfg.IMAGE, cfg.IMAGE)
    y = cfg.traj_labels().cuda()
    assert GraphedStepper.cached(d, lambda x, t, **k: unet(x, t, **k), None, shape, "cuda", {"y": y}, True, False, 0.0) is None
    assert GraphedStepper.cached(d, unet, lambda x, t, **k: x, shape, "cuda", {"y": y}, True, False, 0.0) is None
    assert GraphedStepper.cached(d, unet, None, shape, "cuda", {"y": y, "extra": [1, 2]}, True, False, 0.0) is None
    # the specialised classifier-guided stepper is still chosen where it applies
    st = GraphedStepper.cached(d, ModelFn(unet, True), None, shape, "cuda", {"y": y}, True, False, 0.0)
    assert isinstance(st, GraphedStepper)
