"""Soak test of the step's synchronisation protocols (mbarrier rings, transform warps, producer hand-offs): the SAME
guided step (same x_t, t, labels, noise) replayed many times must give the same BITS every time — a race in the
operand pipeline of the fused convolution would show up as a rare mismatch.

  python profiles/soak_determinism.py [batch] [replays]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th  # noqa: E402

import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 200
dev = th.device("cuda", 0)
th.manual_seed(0)
diffusion, model_fn, cond_fn = bench.build_cfg2(256, dev)
x = th.randn(B, 3, 256, 256, device=dev)
z = th.randn(B, 3, 256, 256, device=dev)
y = th.randint(0, 1000, (B,), device=dev)
t = th.full((B,), 120, dtype=th.int64, device=dev)
ref, bad = None, 0
with th.no_grad():
    for i in range(N):
        out = diffusion._sample_step(model_fn, x, t, True, None, cond_fn, {"y": y}, False, 0.0, noise=z)
        s = out["sample"]
        if ref is None:
            ref = s.clone()
        elif not th.equal(s, ref):
            bad += 1
            print(f"replay {i}: MISMATCH max |diff| {float((s - ref).abs().max()):.3e}", flush=True)
th.cuda.synchronize()
print(f"batch {B}: {N} replays of one guided step, {bad} mismatches, finite={bool(th.isfinite(ref).all())}")
sys.exit(1 if bad else 0)
