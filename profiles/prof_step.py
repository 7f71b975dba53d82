"""One eager (un-graphed) guided step of BASELINE configs[1] (256x256, batch 8) inside an NVTX range "timed", for
ncu launch lists and full captures:

  python profiles/prof_step.py > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv \
      --log-file gpurun_out/launches.csv python profiles/prof_step.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["GD_B200_NO_GRAPH"] = "1"

import torch as th  # noqa: E402

import bench  # noqa: E402
from guided_diffusion_clip_b200 import script_util as su  # noqa: E402
from guided_diffusion_clip_b200.sampler import ClassifierGuidance, ModelFn  # noqa: E402

B = int(os.environ.get("PROF_BATCH", "8"))
S = int(os.environ.get("PROF_SIZE", "256"))
dev = th.device("cuda", 0)
th.manual_seed(0)
model, diffusion = su.create_model_and_diffusion(**bench.unet_kwargs(S))
bench.randomize_(model, 1234)
model.to(dev)
model.convert_to_fp16()
clf = su.create_classifier(**bench.clf_kwargs(S))
bench.randomize_(clf, 4321)
clf.to(dev)
clf.convert_to_fp16()
cond = ClassifierGuidance(clf, 1.0)
mf = ModelFn(model, True)
y = th.randint(0, 1000, (B,), device=dev)
x = th.randn(B, 3, S, S, device=dev)
t = th.full((B,), 200, dtype=th.int64, device=dev)
with th.no_grad():
    for _ in range(2):
        out = diffusion.p_sample(mf, x, t, cond_fn=cond, model_kwargs={"y": y})
    th.cuda.synchronize()
    th.cuda.nvtx.range_push("timed")
    out = diffusion.p_sample(mf, x, t, cond_fn=cond, model_kwargs={"y": y})
    th.cuda.synchronize()
    th.cuda.nvtx.range_pop()
print("ok", float(out["sample"].abs().mean()))
