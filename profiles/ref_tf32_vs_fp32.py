"""How far is the REFERENCE'S OWN GPU arithmetic from its fp32 result?  The reference (unmodified modules from the
git-ignored baseline/_ref) runs BASELINE configs[4] (`use_fp16 False`: UNet-512 + classifier-512, one guided DDIM step,
batch 1, the seeds of tests/golden/fullsize_cfg5.npz) on the B200 through PyTorch/cuDNN twice — with PyTorch's default
TF32 convolutions (torch.backends.cudnn.allow_tf32 = True, what `use_fp16=False` gives a user of the reference on any
Ampere-or-later GPU) and with TF32 switched off — and both are compared with the committed CPU-fp32 fixture, next to
this repository's fp16-storage CUDA path on the same inputs.

  python profiles/ref_tf32_vs_fp32.py [cfg5|cfg2]   ->  one JSON line"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import numpy as np  # noqa: E402
import torch as th  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from guided_diffusion import script_util as rsu  # noqa: E402  (the reference)
from guided_diffusion import unet as runet  # noqa: E402
from oracle import golden_cfg as cfg  # noqa: E402
from oracle.oracle_models import make_state_dict  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
c = cfg.FULLSIZE_CASES[name]
z = np.load(os.path.join(ROOT, "tests", "golden", f"fullsize_{name}.npz"))
G = {k: th.from_numpy(cfg.fs_unpack(z[k], z[k + "_exp"])) for k in ("eps", "grad", "sample")}
dev = th.device("cuda", 0)
seed = cfg.FS_SEED + c["k"]
unet = runet.UNetModel(**(cfg.ref_unet512_kwargs() if name == "cfg5" else dict(cfg.ref_unet256_kwargs(), use_fp16=False)))
unet.load_state_dict(make_state_dict({k: tuple(v.shape) for k, v in unet.state_dict().items()}, seed), strict=True)
clf = rsu.create_classifier(**(cfg.CLF512_KW if name == "cfg5" else cfg.CLF256_KW))
clf.load_state_dict(make_state_dict({k: tuple(v.shape) for k, v in clf.state_dict().items()}, seed + 100), strict=True)
unet.to(dev).eval()
clf.to(dev).eval()
x = cfg.fullsize_inputs(name)[0].to(dev)
t = th.tensor([int(z["t_model"])], device=dev)
y = th.tensor([c["label"]], device=dev)


def rel(a, b):
    return float((a.cpu() - b).abs().max() / b.abs().max())


def run():
    with th.no_grad():
        eps = unet(x, t, y)
    with th.enable_grad():
        xin = x.detach().requires_grad_(True)
        lp = F.log_softmax(clf(xin, t), dim=-1)
        grad = th.autograd.grad(lp[range(1), y.view(-1)].sum(), xin)[0] * c["scale"]
    return eps, grad


out = {"case": name}
for tf32 in (True, False):
    th.backends.cudnn.allow_tf32 = tf32
    th.backends.cuda.matmul.allow_tf32 = False     # PyTorch's defaults: TF32 for cuDNN convolutions only
    eps, grad = run()
    out["reference_gpu_tf32_convs" if tf32 else "reference_gpu_fp32"] = {"eps": rel(eps, G["eps"]), "grad": rel(grad, G["grad"])}
print(json.dumps(out))
