"""Diagnostic (not pytest): compare the autograd-closure and fused guidance paths of the tiny classifier."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch as th
import torch.nn.functional as F

from guided_diffusion_clip_b200 import script_util as su
from guided_diffusion_clip_b200.sampler import ClassifierGuidance
from oracle import golden_cfg as cfg
from oracle import oracle_models as om

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "models_golden.npz"))
clf = su.create_classifier(**cfg.CLASSIFIER_KW)
sd = om.make_state_dict({k: tuple(v.shape) for k, v in clf.state_dict().items()}, cfg.CLF_SEED)
clf.load_state_dict(sd)
clf.cuda().eval()
x, t, y = (v.cuda() for v in cfg.model_inputs())
ref = th.from_numpy(G["clf_grad"]).cuda()


def stats(name, g):
    print(f"{name}: absmax {float(g.abs().max()):.4e} mean {float(g.mean()):.4e} nan {bool(th.isnan(g).any())} "
          f"err {float((g - ref).abs().max() / ref.abs().max()):.3e} ratio {float((g * ref).sum() / (ref * ref).sum()):.4f}")


plan = clf.plan(2, 64, 64, x.device)
for rep in range(2):
    x_in = x.detach().requires_grad_(True)
    with th.enable_grad():
        logits = clf(x_in, t)
        sel = F.log_softmax(logits, -1)[range(2), y]
        g = th.autograd.grad(sel.sum(), x_in)[0]
    stats(f"closure[{rep}]", g)
    dl_torch = plan.dlogits.clone()
    g2 = ClassifierGuidance(clf, 1.0)(x, t, y=y)
    stats(f"fused[{rep}]", g2)
    dl_mine = plan.dlogits.clone()
    print("  dlogits torch absmax", float(dl_torch.abs().max()), "mine", float(dl_mine.abs().max()), "diff",
          float((dl_torch - dl_mine).abs().max()))
    g3 = plan.backward(dl_torch).clone()
    stats(f"bwd-again[{rep}]", g3)
    g4 = plan.backward(dl_mine).clone()
    stats(f"bwd-mine[{rep}]", g4)
