"""Per-call timing of the recorded step programs (UNet fwd, classifier fwd, classifier dX bwd): every recorded
C-ABI call is launched alone `reps` times between CUDA events (warm L2 between reps of the same call — compare
SHARES and TFLOP/s, the step itself is timed by bench.py).  Convs are reported with geometry and TFLOP/s.

  PROF_BATCH=64 python profiles/per_call.py [--top 40]
"""
import argparse
import collections
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th  # noqa: E402

import bench  # noqa: E402
from guided_diffusion_clip_b200 import _lib as L  # noqa: E402
from guided_diffusion_clip_b200.engine import UNetPlan  # noqa: E402


def describe(name, args):
    if name == "gd_conv_igemm":
        d = args[0]._obj if hasattr(args[0], "_obj") else args[0]
        k = d.k_total
        flops = 2.0 * d.n * d.h * d.w * d.cout * k
        tag = f"conv {d.h}x{d.w} K={k} (C0={d.c0} taps={d.taps} C1={d.c1}) -> {d.cout} res={d.res_mode} out={d.out_mode}" \
              f"{' +stats' if d.stats_out else ''}{' +gn%d' % d.gn_mode if d.gn_mode else ''}"
        return tag, flops
    def val(a):
        return a.value if hasattr(a, "value") else a
    if name == "gd_groupnorm_apply":
        n, h, w, c, silu, mode = (val(a) for a in args[9:15])
        aux = bool(val(args[15]))
        out_px = {0: 1.0, 1: 0.25, 2: 4.0}.get(mode, 1.0)  # GD_GN_SAME / AVGPOOL2 / UPSAMPLE2
        nbytes = n * h * w * c * 2.0 * (1.0 + out_px + (0.25 if aux else 0.0))
        return f"gn_apply {h}x{w} C={c} mode={mode} silu={silu} film={bool(val(args[5]))} aux={aux}", -nbytes
    if name == "gd_groupnorm_bwd":
        n, h, w, c, silu, mode = (val(a) for a in args[15:21])
        out_px = {0: 1.0, 1: 0.25, 2: 4.0}.get(mode, 1.0)
        has_add = bool(val(args[9]))
        # two passes read x and dy; second pass writes dx (+ reads add)
        nbytes = n * h * w * c * 2.0 * (2.0 + 2.0 * out_px + 1.0 + (1.0 if has_add else 0.0))
        return f"gn_bwd {h}x{w} C={c} mode={mode} add={has_add}", -nbytes
    return name, 0.0


def time_program(prog, reps):
    stream = C.c_void_p(th.cuda.current_stream().cuda_stream)
    rows = []
    for fn, args, name in prog.calls:
        evs = []
        for _ in range(reps + 1):
            e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args, stream)
            e1.record()
            if rc != 0:
                L.check(rc, name)
            evs.append((e0, e1))
        th.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in evs[1:])
        tag, flops = describe(name, args)
        rows.append((tag, ms[len(ms) // 2], flops))
    return rows


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    batch = int(os.environ.get("PROF_BATCH", "64"))
    dev = th.device("cuda:0")
    from guided_diffusion_clip_b200 import script_util as su
    model, diffusion = su.create_model_and_diffusion(**bench.unet_kwargs(256))
    bench.randomize_(model, 1234)
    model.to(dev)
    model.convert_to_fp16()
    model.eval()
    classifier = su.create_classifier(**bench.clf_kwargs(256))
    bench.randomize_(classifier, 4321)
    classifier.to(dev)
    classifier.convert_to_fp16()
    classifier.eval()
    unet = UNetPlan(model, batch, 256, 256, dev)
    clf = classifier.plan(batch, 256, 256, dev)
    x = th.randn(batch, 3, 256, 256, device=dev)
    unet.x_in.copy_(x)
    unet.t_in.fill_(500.0)
    clf.x_in.copy_(x)
    clf.t_in.fill_(500.0)
    for phase, prog in (("unet_fwd", unet.prog), ("clf_fwd", clf.fwd), ("clf_bwd", clf.bwd)):
        prog.run()
        th.cuda.synchronize()
        rows = time_program(prog, a.reps)
        tot = sum(r[1] for r in rows)
        agg = collections.OrderedDict()
        for tag, ms, fl in rows:
            e = agg.setdefault(tag, [0, 0.0, 0.0])
            e[0] += 1
            e[1] += ms
            e[2] += fl
        print(f"== {phase}: {len(rows)} calls, {tot:.2f} ms (sum of isolated launches), batch {batch}")
        for tag, (cnt, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: a.top]:
            tf = f" {fl / ms / 1e9:7.0f} TF" if fl > 0 else (f" {-fl / ms / 1e6:7.0f} GB/s" if fl < 0 else "")
            print(f"  {ms:8.3f} ms {100 * ms / tot:5.1f}%  n={cnt:3d}{tf}  {tag}")
        print(json.dumps({"phase": phase, "total_ms": round(tot, 3)}))
