"""SURVEY §8f row 3 — the fork's own use-case against the REAL reference (tests/golden/fork_golden.npz, produced by
oracle/make_golden_fork.py): SRImageModel_Feat.forward (unet_other.py:43-77) and the denoise_start_point /
q_sample(img2) start of p_sample_loop (gaussian_diffusion.py:517-523)."""
import os

import numpy as np
import pytest
import torch as th

from guided_diffusion_clip_b200 import script_util as su
from oracle import golden_cfg as cfg
from oracle import oracle_models as om
from tests import gpu_helpers as H

TOL = 2e-2


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "fork_golden.npz"))


def _model():
    m = su.sr_create_model(**{k: v for k, v in cfg.SRFEAT_KW.items()
                              if k not in ("timestep_respacing",)}, conditioning="clip_feat")
    sd = om.make_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, cfg.SRFEAT_SEED)
    m.load_state_dict(sd, strict=True)
    return m, sd


def test_oracle_srfeat_matches_reference(G):
    """CPU: the oracle restatement of SRImageModel_Feat against the reference's output."""
    _, sd = _model()
    x, t, f1, f2, img2 = cfg.srfeat_inputs()
    with th.no_grad():
        out = om.srfeat_forward(sd, x, t, f1, f2, img2, **cfg.SR_STRUCT)
    assert H.rel_err(out, th.from_numpy(G["srfeat_out"])) < 2e-4


def test_srfeat_state_dict_layout_matches_reference_keys():
    """bias_feat [512] + label MLP + 6-channel first conv (SURVEY App. E 'fork extras')."""
    m, sd = _model()
    assert tuple(sd["bias_feat"].shape) == (512,)
    assert tuple(sd["label_emb.0.weight"].shape) == (256, 512) and tuple(sd["label_emb.2.weight"].shape) == (256, 256)
    assert tuple(sd["input_blocks.0.0.weight"].shape) == (64, 6, 3, 3)


@pytest.mark.gpu
def test_srfeat_forward_matches_reference(lib, G):
    m, _ = _model()
    m.cuda().eval()
    x, t, f1, f2, img2 = (v.cuda() for v in cfg.srfeat_inputs())
    with th.no_grad():
        out = m(x, t, clip_feat=f1, clip_feat2=f2, img2=img2)
    err = H.rel_err(out, th.from_numpy(G["srfeat_out"]).cuda())
    print(f"SRImageModel_Feat vs reference golden: rel err {err:.3e}")
    assert out.shape == (2, 6, 64, 64) and err < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("graph", ["0", "1"])
def test_denoise_start_point_loop_matches_reference(lib, G, monkeypatch, graph):
    """p_sample_loop(..., denoise_start_point=40) on a 250-step chain: starts at q_sample(img2, t=40), runs steps 39..0.
    The reference drew its noise from the CPU generator (randn(shape), q_sample's randn_like, one randn_like per step);
    the same draws are replayed here by routing torch.randn / randn_like / Tensor.normal_ through the CPU generator."""
    monkeypatch.setenv("GD_B200_NO_GRAPH", "0" if graph == "1" else "1")
    m, _ = _model()
    m.cuda().eval()
    x, _, f1, f2, img2 = (v.cuda() for v in cfg.srfeat_inputs())
    d = su.create_gaussian_diffusion(**cfg.SRFEAT_DIFFUSION)
    real_randn = th.randn

    def cpu_randn(*shape, device=None, **kw):
        shape = shape[0] if len(shape) == 1 and isinstance(shape[0], (tuple, list, th.Size)) else shape
        return real_randn(*shape).to(device if device is not None else "cpu")

    monkeypatch.setattr(th, "randn", cpu_randn)
    monkeypatch.setattr(th, "randn_like", lambda v: real_randn(*v.shape).to(v.device))
    monkeypatch.setattr(th.Tensor, "normal_", lambda self: self.copy_(real_randn(*self.shape)))
    seen = []

    class Spy:
        def __init__(self, inner):
            self.inner = inner

        def __call__(self, x_, t_, **kw):
            seen.append(int(t_[0]))
            return self.inner(x_, t_, **kw)

        def parameters(self):
            return self.inner.parameters()

    model = m if graph == "1" else Spy(m)
    th.manual_seed(cfg.SRFEAT_LOOP_SEED)
    mk = {"clip_feat": f1, "clip_feat2": f2, "img2": img2}
    steps = [o["sample"] for o in d.p_sample_loop_progressive(model, tuple(x.shape), model_kwargs=mk, device=x.device,
                                                               denoise_start_point=cfg.SRFEAT_START)]
    ref = th.from_numpy(G["srfeat_loop"]).cuda()
    assert len(steps) == cfg.SRFEAT_START
    if graph == "0":
        assert seen == [int(v) for v in G["srfeat_loop_ts"]]  # the original timesteps 156, 152, ... 0 reach the model
    errs = [H.rel_err(steps[k - 1], r) for k, r in zip(cfg.SRFEAT_RECORD, ref)]
    print(f"denoise_start_point loop (graph={graph}) vs reference: per-step rel err " + " ".join(f"{e:.3e}" for e in errs))
    assert max(errs) < TOL
    th.manual_seed(cfg.SRFEAT_LOOP_SEED)
    final = d.p_sample_loop(model, tuple(x.shape), model_kwargs=mk, device=x.device,
                            denoise_start_point=cfg.SRFEAT_START)
    assert th.equal(final, steps[-1])
