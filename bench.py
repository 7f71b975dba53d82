#!/usr/bin/env python
"""bench.py — guided-diffusion sampling on N B200s of one node (BASELINE.json: headline = configs[1]).

    python bench.py --gpus N --steps K --warmup W                  # this repo's CUDA path, configs[1]
    python bench.py --impl reference --steps K --warmup W          # the reference's own CPU implementation (host cores)
    python bench.py --config clip256|sr512|adm512 [--batch B]      # configs[2..4] through the same harness

A "step" is one guided sampling step on one per-GPU batch: UNet forward + guidance forward + guidance data-gradient
backward + fused posterior/noise update (gaussian_diffusion.py:395-439 / 546-594 of the reference).  A configs[1]
sample needs 250 such steps (timestep_respacing="250"), so  value [samples/s] = n_gpus * batch / (250 * s_per_step).

What one line holds (all measured in this process, nothing quoted):
  value / ms_per_step  K guided steps at 64 samples per GPU ("scaling": "weak"), CUDA events, max over ranks
  strong               BASELINE's own split: GLOBAL batch 64 -> 64/N samples per GPU, same timing method
  full_loop            ONE whole 250-step diffusion.p_sample_loop at the weak batch + uint8-NHWC pack into the gather
                       buffer + the path's only collective (all_gather of samples and labels), device-timed
  e2e                  the step through the public API with pinned HOST buffers (H2D of x_t and t, D2H of x_{t-1})
  roofline             the dominant kernel (3x3 conv 256->256 @256x256 with the GroupNorm+SiLU operand transform fused)
                       timed back to back in the hot, power-capped state against the SUSTAINED measured tensor peak,
                       plus the same kernel alone against the burst peak and the whole step's TFLOP/s;
                       DRAM traffic / tensor-pipe activity are read from the committed ncu summary (profiles/)
  cpu_baseline         the reference's own modules (vendored at build time into git-ignored baseline/_ref) on the host
                       cores, bounded sample: guided steps at batch 1, extrapolated x250 and labelled so
Synthetic data: N(0,1) noise of the named shape, random-init weights (zero_module tensors re-drawn N(0,0.02)).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch as th  # noqa: E402

STEPS_PER_SAMPLE = 250
GFLOP_PER_SAMPLE_STEP = 2535.5  # UNet-256 2239.67 + classifier fwd 146.70 + dX bwd 149.15
NCU_SUMMARY = os.path.join(ROOT, "profiles", "ncu_summary_r02.json")
REF_DIR = os.path.join(ROOT, "baseline", "_ref")

# configs of BASELINE.json beyond the headline: (steps per sample, algorithmic GFLOP per sample and step, default batch)
OTHER_CONFIGS = {
    "clip256": dict(index=2, steps=50, gflop=2239.67 + 3 * 35.1, batch=32, image=256,
                    workload="256x256 unconditional ADM + CLIP ViT-B/16 image-encoder guidance (random-init CLIP, "
                             "scale 100), DDIM 50 steps"),
    "sr512": dict(index=3, steps=250, gflop=5009.57, batch=8, image=512,
                  workload="128->512 super-res upsampler (192 ch, attn 32/16), synthetic low-res conditioning, 250 steps"),
    "adm512": dict(index=4, steps=25, gflop=3964.67 + 2 * 225.53, batch=8, image=512,
                   workload="512x512 class-cond ADM (use_fp16 False masters) + classifier-512 guidance (scale 4.0), "
                            "DDIM 25 steps"),
}


def unet_kwargs(image_size=256):
    from guided_diffusion_clip_b200 import script_util as su
    d = su.model_and_diffusion_defaults()
    d.update(image_size=image_size, num_channels=256, num_res_blocks=2, attention_resolutions="32,16,8",
             num_head_channels=64, resblock_updown=True, use_scale_shift_norm=True, learn_sigma=True, class_cond=True,
             use_fp16=True, noise_schedule="linear", diffusion_steps=1000, timestep_respacing="250")
    return d


def clf_kwargs(image_size=256):
    from guided_diffusion_clip_b200 import script_util as su
    d = su.classifier_defaults()
    d.update(image_size=image_size, classifier_use_fp16=True)
    return d


def workload_config(image_size, batch, world, note=""):
    """The `config` object shared by both arms (the reference arm states its own bounded sample in `sample`)."""
    return {"workload": f"{image_size}x{image_size} class-cond ADM (256ch, 2 res blocks, attn 32/16/8) + EncoderUNet "
                        f"classifier guidance (scale 1.0), 250 respaced steps, batch {batch}/GPU (global {batch * world}), "
                        "fp16 storage fp32 accumulate" + note,
            "per_gpu_batch": batch, "global_batch": batch * world, "steps_per_sample": STEPS_PER_SAMPLE,
            "l2_policy": "activations per step (>1 GB at batch 8) exceed the 126 MB L2; no flush needed"}


def randomize_(model, seed):
    """Random init; every all-zero (zero_module) tensor re-drawn N(0, 0.02) so no branch is vacuous."""
    g = th.Generator(device="cpu").manual_seed(seed)
    with th.no_grad():
        for p in model.parameters():
            if float(p.abs().max()) == 0.0:
                p.copy_(th.randn(p.shape, generator=g) * 0.02)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons, power = [], 0.0, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                power.append(float(r[2]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "power_w": statistics.median(power) if power else None, "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


def ncu_entry(key):
    """Evidence captured with `ncu --set full` on the final build of the round, summarised by profiles/ncu_summarize.py
    into a committed JSON file; None if the file or the key is absent (then the field is reported as null)."""
    try:
        with open(NCU_SUMMARY) as f:
            return json.load(f).get(key)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------------------
# CPU arm
# ------------------------------------------------------------------------------------------------------------
def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "guided_diffusion", "unet.py"))


def cpu_guided_steps_reference(n_steps: int, image_size: int, batch: int = 1):
    """`n_steps` guided p_sample steps of the REAL reference (its unmodified modules, vendored by __graft_entry__.build()
    into git-ignored baseline/_ref): unet.UNetModel + create_classifier + create_gaussian_diffusion in fp32 with the
    cond_fn / model_fn closures of scripts/classifier_sample.py:54-65, all host threads.  Seconds per step."""
    import torch.nn.functional as F
    sys.path.insert(0, REF_DIR)
    from guided_diffusion import script_util as rsu  # noqa: E402  (the reference)
    from guided_diffusion import unet as runet  # noqa: E402
    th.set_num_threads(os.cpu_count() or 1)
    mult = {256: (1, 1, 2, 2, 4, 4), 512: (0.5, 1, 1, 2, 2, 4, 4)}[image_size]
    th.manual_seed(0)
    model = runet.UNetModel(image_size=image_size, in_channels=3, model_channels=256, out_channels=6, num_res_blocks=2,
                            attention_resolutions=tuple(image_size // r for r in (32, 16, 8)), dropout=0.0,
                            channel_mult=mult, num_classes=1000, use_checkpoint=False, use_fp16=False, num_heads=4,
                            num_head_channels=64, num_heads_upsample=-1, use_scale_shift_norm=True, resblock_updown=True,
                            use_new_attention_order=False).eval()
    ckw = rsu.classifier_defaults()
    ckw.update(image_size=image_size)
    classifier = rsu.create_classifier(**ckw).eval()
    randomize_(model, 1234)
    randomize_(classifier, 4321)
    diffusion = rsu.create_gaussian_diffusion(steps=1000, learn_sigma=True, noise_schedule="linear",
                                              timestep_respacing="250")

    def cond_fn(x, t, y=None):
        with th.enable_grad():
            x_in = x.detach().requires_grad_(True)
            logits = classifier(x_in, t)
            log_probs = F.log_softmax(logits, dim=-1)
            selected = log_probs[range(len(logits)), y.view(-1)]
            return th.autograd.grad(selected.sum(), x_in)[0] * 1.0

    def model_fn(x, t, y=None):
        return model(x, t, y)

    x = th.randn(batch, 3, image_size, image_size)
    y = th.randint(0, 1000, (batch,))
    times = []
    for s in range(n_steps):
        t = th.tensor([diffusion.num_timesteps - 1 - s] * batch)
        t0 = time.perf_counter()
        with th.no_grad():
            x = diffusion.p_sample(model_fn, x, t, cond_fn=cond_fn, model_kwargs={"y": y})["sample"]
        times.append(time.perf_counter() - t0)
    return times


def cpu_guided_steps_port(n_steps: int, image_size: int, batch: int = 1):
    """The same bounded sample through the oracle port (oracle/, a CPU restatement pinned to the reference by
    tests/golden) — the fallback when baseline/_ref is absent."""
    from guided_diffusion_clip_b200 import script_util as su
    from oracle import oracle_diffusion as od
    from oracle import oracle_models as om
    th.set_num_threads(os.cpu_count() or 1)
    with th.device("meta"):
        um = su.create_model(**{k: v for k, v in unet_kwargs(image_size).items()
                                if k in su.create_model.__code__.co_varnames})
        cm = su.create_classifier(**clf_kwargs(image_size))
    usd = om.make_state_dict({k: tuple(v.shape) for k, v in um.state_dict().items()}, 1)
    csd = om.make_state_dict({k: tuple(v.shape) for k, v in cm.state_dict().items()}, 2)
    tab = od.Tables(schedule="linear", steps=1000, respacing="250", learn_sigma=True)
    ustruct = dict(num_res_blocks=2, channel_mult_len=len(um.channel_mult), head_dim=64, new_order=False)
    cstruct = dict(num_res_blocks=2, channel_mult_len=len(cm.channel_mult), head_dim=64)
    g = th.Generator().manual_seed(0)
    x = th.randn(batch, 3, image_size, image_size, generator=g)
    y = th.randint(0, 1000, (batch,), generator=g)
    times = []
    for s in range(n_steps):
        i = tab.T - 1 - s
        t0 = time.perf_counter()
        tt = th.full((batch,), tab.timestep_map[i])
        with th.no_grad():
            mo = om.unet_forward(usd, x, tt, y, **ustruct)
        z = th.randn(x.shape, generator=g)
        grad = om.classifier_guidance(csd, x, tt, y, 1.0, **cstruct)
        x = tab.p_sample(mo, x, i, z, grad)["sample"]
        times.append(time.perf_counter() - t0)
    return times


def cpu_guided_steps(n_steps, image_size, batch=1):
    """(seconds per step list, kind): the real reference when it was vendored, else the port."""
    if reference_available() and os.environ.get("GD_B200_CPU_PORT", "0") != "1":
        try:
            return cpu_guided_steps_reference(n_steps, image_size, batch), "reference"
        except Exception as e:  # noqa: BLE001
            sys.stderr.write(f"bench.py: vendored reference failed ({e!r}); timing the oracle port instead\n")
    return cpu_guided_steps_port(n_steps, image_size, batch), "port"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times, kind = cpu_guided_steps(args.warmup + args.steps, args.image_size, 1)
    timed = times[args.warmup:]
    s_per_step = sum(timed) / len(timed)
    value = 1.0 / (STEPS_PER_SAMPLE * s_per_step)
    cores = os.cpu_count() or 1
    what = ("the reference's own modules (baseline/_ref, unmodified)" if kind == "reference" else
            "the oracle port of the reference algorithm (oracle/)")
    sample = (f"{len(timed)} guided p_sample steps at batch 1 (the GPU arm runs batch 64/GPU), {args.image_size}x"
              f"{args.image_size}, fp32 oneDNN, {cores} threads, {what}; samples/s extrapolated x{STEPS_PER_SAMPLE} steps")
    line = {
        "impl": "reference", "metric": "guided_samples_per_sec_256", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.image_size, 1, 1, " [CPU arm: bounded sample at batch 1 on ONE host, whatever N]"),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------------------------
# GPU arm: kernel-level roofline measurements
# ------------------------------------------------------------------------------------------------------------
def top_conv_roofline(batch, image_size, burst_tflops, sustained_tflops, hot=True):
    """The dominant kernel: 3x3 conv 256->256 at full resolution WITH the GroupNorm+FiLM+SiLU operand transform fused
    (conv_igemm_kernel<pair, gn>; 31 % of the step's FLOPs).  Two measurements with CUDA events on the
    launching stream: `sustained` = 48 launches back to back right after the timed steps (hot, power-capped — the state
    the kernel runs in inside a step; denominator = the sustained measured peak) and `burst` = 8 separate launches
    (denominator = the burst peak).  Inputs (268 MB at batch 8) exceed the 126 MB L2; launches rotate over two buffers."""
    import ctypes as C
    from guided_diffusion_clip_b200 import _lib as L
    from guided_diffusion_clip_b200.engine import pack_conv3x3
    lib = L.load()
    c = 256
    xs = [th.randn((batch, image_size, image_size, c), device="cuda", dtype=th.float16) for _ in range(2)]
    out = th.empty_like(xs[0])
    w = pack_conv3x3(th.randn((c, c, 3, 3), device="cuda") * 0.02)
    bias = th.zeros(c, device="cuda")
    coef = th.zeros((batch, c // 8, 16), device="cuda")
    coef[..., :8] = 1.0   # identity affine: y = SiLU(x)
    d = L.ConvDesc()
    d.c0, d.ld0, d.taps, d.n, d.h, d.w = c, c, 9, batch, image_size, image_size
    d.wpack, d.k_total, d.n_pad, d.bias, d.cout = w.data_ptr(), 9 * c, c, bias.data_ptr(), c
    d.out, d.ld_out, d.out_mode, d.out_scale = out.data_ptr(), c, L.OUT_NHWC_F16, 1.0
    d.gn_mode, d.gn_silu, d.gn_coef = L.CONV_GN_SAME, 1, coef.data_ptr()
    stream = C.c_void_p(th.cuda.current_stream().cuda_stream)

    def launch(i):
        d.a0 = xs[i % 2].data_ptr()
        L.check(lib.gd_conv_igemm(C.byref(d), stream), "gd_conv_igemm")

    flops = 2.0 * batch * image_size * image_size * c * 9 * c
    n_hot = 48
    launch(0)
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_hot):
        launch(i)
    e1.record()
    th.cuda.synchronize()
    ms_hot = e0.elapsed_time(e1) / n_hot
    evs = []
    for i in range(8):
        a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        a.record()
        launch(i)
        b.record()
        th.cuda.synchronize()
        evs.append(a.elapsed_time(b))
        time.sleep(0.05)
    ms_burst = sum(evs) / len(evs)
    ach_hot = flops / (ms_hot * 1e-3) / 1e12
    ach_burst = flops / (ms_burst * 1e-3) / 1e12
    algo_bytes = 2.0 * batch * image_size * image_size * c * 2 + 9 * c * c * 2
    ev = ncu_entry("conv_gn_256_256_b64") or {}
    traffic = ev.get("dram_bytes_per_launch")
    if traffic is not None and ev.get("batch"):
        traffic = traffic * batch / ev["batch"]
    return {"bound": "tensor", "kernel": "conv_igemm_kernel<pair,gn> 3x3 256->256 @%dx%d batch %d, GroupNorm+FiLM+SiLU "
                                         "fused into the operand path" % (image_size, image_size, batch),
            "achieved": ach_hot, "peak": sustained_tflops, "unit": "TFLOP/s", "frac": ach_hot / sustained_tflops,
            "peak_kind": "sustained (kernel timed back to back in the power-capped state)",
            "avg_launch_ms": ms_hot, "flops_per_launch": flops, "launches_timed": n_hot,
            "burst": {"achieved": ach_burst, "peak": burst_tflops, "frac": ach_burst / burst_tflops,
                      "avg_launch_ms": ms_burst},
            "traffic": traffic, "traffic_unit": "bytes/launch", "algorithmic_bytes": algo_bytes,
            "traffic_source": ev.get("source", "no committed ncu summary for this kernel"),
            "tensor_pipe_active_pct_ncu": ev.get("tensor_pipe_active_pct")}


def hbm_kernels_roofline(diffusion, batch, image_size, hbm_gbs, reps=10):
    """The bandwidth-bound kernels of the step timed alone (CUDA events, current stream), rotating over 3 input sets so
    that no launch finds its inputs in the 126 MB L2: the fused posterior update (21 channels x 4 B per
    pixel) and the classifier's GroupNorm data-gradient — the largest remaining GroupNorm pass now that the forward
    apply is fused into the convs.  Its ALGORITHMIC bytes are x, dy and the residual gradient read once and dx written
    once (4 x 2 B per element); the implementation reads x and dy twice (a statistics pass, then the apply pass), so
    its DRAM traffic is 6 x 2 B per element and `frac` (algorithmic / time / peak) is bounded by 4/6."""
    import ctypes as C
    from guided_diffusion_clip_b200 import _lib as L
    lib = L.load()
    dev = th.device("cuda", th.cuda.current_device())
    shape = (batch, 3, image_size, image_size)
    sets = [dict(x=th.randn(shape, device=dev), mo=th.randn((batch, 6) + shape[2:], device=dev),
                 g=th.randn(shape, device=dev), z=th.randn(shape, device=dev)) for _ in range(3)]
    sample, x0 = th.empty(shape, device=dev), th.empty(shape, device=dev)
    t = th.full((batch,), diffusion.num_timesteps // 2, dtype=th.int64, device=dev)
    out = []

    def timed(fn):
        evs = []
        for i in range(reps + 3):
            e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
            e0.record()
            fn(i)
            e1.record()
            evs.append((e0, e1))
        th.cuda.synchronize()
        ms = [a.elapsed_time(b) for a, b in evs[3:]]
        return sum(ms) / len(ms)

    ms = timed(lambda i: diffusion._launch_posterior(x=sets[i % 3]["x"], t=t, model_out=sets[i % 3]["mo"],
                                                     grad=sets[i % 3]["g"], noise=sets[i % 3]["z"], sample=sample,
                                                     pred_xstart=x0))
    nbytes = 21.0 * 4 * batch * image_size * image_size
    gbs = nbytes / (ms * 1e-3) / 1e9
    ev = ncu_entry("posterior_b64") or {}
    out.append({"kernel": "posterior_kernel (guided p_sample update)", "bound": "hbm", "achieved": gbs, "peak": hbm_gbs,
                "unit": "GB/s", "frac": gbs / hbm_gbs, "avg_launch_ms": ms, "algorithmic_bytes": nbytes,
                "traffic": ev.get("dram_bytes_per_launch"), "traffic_source": ev.get("source")})
    del sets
    c = 128
    xs = [th.randn((batch, image_size, image_size, c), device=dev, dtype=th.float16) for _ in range(3)]
    dys = [th.randn((batch, image_size, image_size, c), device=dev, dtype=th.float16) for _ in range(3)]
    dx = th.empty_like(xs[0])
    st = th.zeros((batch, 32, 2), device=dev)
    st[..., 1] = 1.0
    gamma, beta = th.ones(c, device=dev), th.zeros(c, device=dev)
    ws = th.empty(int(lib.gd_groupnorm_ws_floats(batch, image_size * image_size, c)), device=dev)
    stream = C.c_void_p(th.cuda.current_stream().cuda_stream)
    vp = lambda v: C.c_void_p(v.data_ptr())  # noqa: E731

    def gnb(i):
        L.check(lib.gd_groupnorm_bwd(vp(xs[i % 3]), c, vp(st), vp(gamma), vp(beta), None, 0, vp(dys[i % 3]), c,
                                     vp(dys[(i + 1) % 3]), c, L.GN_SAME, vp(dx), c, vp(ws), batch, image_size, image_size,
                                     c, 1, L.GN_SAME, stream), "gd_groupnorm_bwd")

    ms = timed(gnb)
    nbytes = 4.0 * 2 * batch * image_size * image_size * c
    gbs = nbytes / (ms * 1e-3) / 1e9
    ev = ncu_entry("gn_bwd_128_b64") or {}
    out.append({"kernel": "gn_bwd (GroupNorm32+SiLU data-gradient + residual add, 128 ch @%dx%d, 3 launches)"
                          % (image_size, image_size), "bound": "hbm", "achieved": gbs, "peak": hbm_gbs, "unit": "GB/s",
                "frac": gbs / hbm_gbs, "avg_launch_ms": ms, "algorithmic_bytes": nbytes,
                "moved_bytes_two_pass": 1.5 * nbytes, "moved_gbs": 1.5 * gbs,
                "traffic": ev.get("dram_bytes_per_launch"), "traffic_source": ev.get("source")})
    return out


# ------------------------------------------------------------------------------------------------------------
# GPU arm: workloads
# ------------------------------------------------------------------------------------------------------------
def build_cfg2(S, dev):
    from guided_diffusion_clip_b200 import script_util as su
    from guided_diffusion_clip_b200.sampler import ClassifierGuidance, ModelFn
    model, diffusion = su.create_model_and_diffusion(**unet_kwargs(S))
    randomize_(model, 1234)
    model.to(dev)
    model.convert_to_fp16()
    model.eval()
    classifier = su.create_classifier(**clf_kwargs(S))
    randomize_(classifier, 4321)
    classifier.to(dev)
    classifier.convert_to_fp16()
    classifier.eval()
    return diffusion, ModelFn(model, True), ClassifierGuidance(classifier, 1.0)


def time_steps(diffusion, model_fn, cond_fn, B, S, dev, steps, warmup, world, ddim=False, extra_kwargs=None,
               clocks_index=None):
    """`steps` guided steps at per-GPU batch B through the public per-step API (a cached CUDA-graph replay for our own
    objects), bracketed by barrier + synchronize, CUDA events, MAX over ranks.  Returns (ms_per_step, finite, stepper,
    clocks summary, last sample, labels)."""
    import torch.distributed as dist
    from guided_diffusion_clip_b200.sampler import GraphedStepper
    shape = (B, 3, S, S)
    kwargs = dict(extra_kwargs or {})
    y = None
    m = model_fn.model if hasattr(model_fn, "model") else model_fn
    if getattr(m, "num_classes", None) is not None and not getattr(m, "label_mlp", False):
        y = th.randint(0, 1000, (B,), device=dev)
        kwargs["y"] = y
    stepper = GraphedStepper.cached(diffusion, model_fn, cond_fn, shape, dev, kwargs, True, ddim, 0.0)
    assert stepper is not None, "fast path not taken"
    T = diffusion.num_timesteps
    img = th.randn(*shape, device=dev)
    t = th.empty((B,), dtype=th.int64, device=dev)
    fn = diffusion.ddim_sample if ddim else diffusion.p_sample

    def one_step(k):
        nonlocal img
        t.fill_(T - 1 - (k % T))
        img = fn(model_fn, img, t, cond_fn=cond_fn, model_kwargs=kwargs)["sample"]

    for k in range(warmup):
        one_step(k)
    th.cuda.synchronize()
    if world > 1:
        dist.barrier()
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    with ClockSampler(clocks_index if clocks_index is not None else th.cuda.current_device()) as clocks:
        e0.record()
        for k in range(steps):
            one_step(warmup + k)
        e1.record()
        th.cuda.synchronize()
    if world > 1:
        dist.barrier()
    th.cuda.synchronize()
    ms_total = th.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    return float(ms_total) / steps, bool(th.isfinite(img).all()), stepper, clocks.summary(), img, y, kwargs


def run_gpu_arm(args):
    from guided_diffusion_clip_b200 import dist_util
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not th.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU arm has no CPU fallback (use --impl reference)")
    th.cuda.set_device(local)
    if world > 1:
        dist_util.setup_dist("nccl")
    dev = th.device("cuda", local)
    B, S = args.batch, args.image_size
    if world > 1:
        # warm-up of the collective itself: NCCL builds its communicator (rings, NVLS buffers) lazily inside the first
        # call — a one-time cost of ~1.5 s at 8 ranks that belongs to process start-up, not to a sampling batch
        w_in = th.zeros(16, dtype=th.uint8, device=dev)
        w_out = th.zeros(16 * world, dtype=th.uint8, device=dev)
        for _ in range(2):
            dist.all_gather_into_tensor(w_out, w_in)
        th.cuda.synchronize()
    th.manual_seed(dist_util.rank_seed(args.seed, rank))
    diffusion, model_fn, cond_fn = build_cfg2(S, dev)

    ms_per_step, finite, stepper, clocks, img, y, kwargs = time_steps(
        diffusion, model_fn, cond_fn, B, S, dev, args.steps, args.warmup, world, clocks_index=local)
    launches_per_step = stepper.launches_per_step
    T = diffusion.num_timesteps
    shape = (B, 3, S, S)

    # ---- end-to-end through the public API with HOST buffers: H2D of the step's inputs, D2H of its result -------
    host_x = th.randn(*shape).pin_memory()
    host_t = th.empty((B,), dtype=th.int64).pin_memory()
    host_out = th.empty(shape).pin_memory()
    dx = th.empty(shape, device=dev)
    dt = th.empty((B,), dtype=th.int64, device=dev)

    host = [host_x, host_out]

    def e2e_step(k):
        # x_t comes from pinned host memory and x_{t-1} goes back to pinned host memory every step; the two host buffers
        # swap roles (the result of step k is the input of step k + 1) instead of being copied on the host
        src, dst = host[k & 1], host[(k + 1) & 1]
        host_t.fill_(T - 1 - (k % T))
        dx.copy_(src, non_blocking=True)
        dt.copy_(host_t, non_blocking=True)
        out = diffusion.p_sample(model_fn, dx, dt, cond_fn=cond_fn, model_kwargs=kwargs)
        dst.copy_(out["sample"], non_blocking=True)
        th.cuda.current_stream().synchronize()

    for k in range(3):
        e2e_step(k)
    th.cuda.synchronize()
    n_e2e = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for k in range(n_e2e):
        e2e_step(k)
    th.cuda.synchronize()
    e2e_s = th.tensor([(time.perf_counter() - t0) / n_e2e], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * B / (STEPS_PER_SAMPLE * float(e2e_s))
    bytes_io = B * 3 * S * S * 4

    # ---- the whole job once: 250-step p_sample_loop + uint8 pack into the gather buffer + all_gather -------------
    full = None
    if not args.no_full_loop:
        gbuf = dist_util.GatherBuffer(B, 3, S, S, dev)
        # warm-up of the output stage at its real size (NCCL sets up the channels / protocol of a message size on first
        # use: 0.7 - 1.5 s for the first 50 - 100 MB gather of a process, a one-time cost)
        gbuf.pack_and_gather(th.zeros(shape, device=dev), y)
        if world > 1:
            dist.barrier()
        th.cuda.synchronize()
        f0, fl, f1, f2 = (th.cuda.Event(enable_timing=True) for _ in range(4))
        f0.record()
        with th.no_grad():
            final = diffusion.p_sample_loop(model_fn, shape, cond_fn=cond_fn, model_kwargs=kwargs, device=dev)
        fl.record()
        if world > 1:
            # the gather is a rendezvous: a rank that finished its 250 steps early waits here for the slowest one.  The
            # barrier separates that wait (rank_skew_ms) from the cost of the output stage itself (gather_ms); the
            # whole-job time f0 -> f2 contains both, as it does for the reference's driver.
            dist.barrier()
        f1.record()
        imgs, labs = gbuf.pack_and_gather(final, y)
        f2.record()
        th.cuda.synchronize()
        if world > 1:
            dist.barrier()
        loop_ms = th.tensor([f0.elapsed_time(f2), f1.elapsed_time(f2), fl.elapsed_time(f1)], device=dev)
        if world > 1:
            dist.all_reduce(loop_ms, op=dist.ReduceOp.MAX)
        ok = bool(th.isfinite(final).all()) and len(imgs) == world and imgs[0].dtype == th.uint8
        full = {"steps": T, "seconds": float(loop_ms[0]) / 1e3, "value": world * B / (float(loop_ms[0]) / 1e3),
                "unit": "samples/s", "ms_per_step": float(loop_ms[0]) / T, "gather_ms": float(loop_ms[1]),
                "rank_skew_ms": float(loop_ms[2]),
                "gathered_samples": int(sum(int(i.shape[0]) for i in imgs)), "ok": ok,
                "what": "diffusion.p_sample_loop (250 steps) + uint8 NHWC pack into the gather buffer + all_gather of "
                        "samples and labels, device-timed, max over ranks"}
        del gbuf, final, imgs, labs

    # ---- BASELINE's own split: GLOBAL batch 64 -> 64 / N per GPU --------------------------------------------------
    strong = None
    gb = args.strong_global_batch
    if not args.no_strong and gb % world == 0 and gb // world >= 1:
        bs = gb // world
        if bs == B:
            strong = {"global_batch": gb, "per_gpu_batch": bs, "ms_per_step": ms_per_step,
                      "value": gb / (STEPS_PER_SAMPLE * ms_per_step * 1e-3), "unit": "samples/s", "same_as_weak": True}
        else:
            diffusion.__dict__.pop("_steppers", None)  # drop the batch-64 graph and its buffers first
            del stepper
            th.cuda.empty_cache()
            ms_s, fin_s, st_s, clk_s, _, _, _ = time_steps(diffusion, model_fn, cond_fn, bs, S, dev, max(args.steps, 20),
                                                          max(args.warmup, 5), world, clocks_index=local)
            strong = {"global_batch": gb, "per_gpu_batch": bs, "ms_per_step": ms_s,
                      "value": gb / (STEPS_PER_SAMPLE * ms_s * 1e-3), "unit": "samples/s", "finite": fin_s,
                      "launches_per_step": st_s.launches_per_step, "sm_mhz": clk_s.get("sm_mhz"),
                      "step_tflops_per_gpu": GFLOP_PER_SAMPLE_STEP * 1e9 * bs / (ms_s * 1e-3) / 1e12}

    if rank != 0:
        return
    burst, sustained, hbm, src = peaks()
    roof = top_conv_roofline(B, S, burst, sustained)
    roof["peak_source"] = src
    step_tflops = GFLOP_PER_SAMPLE_STEP * 1e9 * B / (ms_per_step * 1e-3) / 1e12
    roof["step_tflops"] = step_tflops
    roof["step_frac_of_sustained"] = step_tflops / sustained
    roof["step_frac_of_burst"] = step_tflops / burst
    roof_hbm = hbm_kernels_roofline(diffusion, B, S, hbm)
    value = world * B / (STEPS_PER_SAMPLE * ms_per_step * 1e-3)
    line = {
        "metric": "guided_samples_per_sec_256", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": workload_config(S, B, world), "finite": finite,
        "roofline": roof, "roofline_hbm": roof_hbm, "clocks": clocks, "gpu_launches": launches_per_step * args.steps,
        "launches_per_step": launches_per_step, "strong": strong, "full_loop": full,
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": bytes_io + B * 8,
                "d2h_bytes_per_step": bytes_io, "steps": n_e2e},
    }
    if world == 1 and not args.no_cpu_baseline:
        times, kind = cpu_guided_steps(1 + args.cpu_steps, S, 1)
        s = sum(times[1:]) / len(times[1:])
        cores = os.cpu_count() or 1
        line["cpu_baseline"] = {
            "value": 1.0 / (STEPS_PER_SAMPLE * s), "unit": "samples/s", "cores": cores, "kind": kind,
            "sample": f"{args.cpu_steps} guided step(s) at batch 1 after 1 warm-up, {S}x{S}, fp32, {cores} threads, "
                      + ("the reference's own modules from baseline/_ref" if kind == "reference" else "oracle port")
                      + f"; extrapolated x{STEPS_PER_SAMPLE} steps ({s:.2f} s/step)"}
    emit(line)


def run_other_config(args):
    """configs[2..4] of BASELINE.json through the same harness (single-GPU bench lines; weak replicas under torchrun)."""
    import torch.nn.functional as F
    from guided_diffusion_clip_b200 import dist_util, script_util as su
    from guided_diffusion_clip_b200.sampler import ClassifierGuidance, ModelFn
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not th.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU arm has no CPU fallback")
    th.cuda.set_device(local)
    if world > 1:
        dist_util.setup_dist("nccl")
    dev = th.device("cuda", local)
    c = OTHER_CONFIGS[args.config]
    B, S = (args.batch if args.batch_given else c["batch"]), c["image"]
    th.manual_seed(dist_util.rank_seed(args.seed, rank))
    extra, ddim, dtype = {}, False, "f16"
    if args.config == "clip256":
        from guided_diffusion_clip_b200 import clip as gclip
        kw = unet_kwargs(256)
        kw.update(class_cond=False, timestep_respacing="ddim50")
        model, diffusion = su.create_model_and_diffusion(**kw)
        randomize_(model, 11)
        model.to(dev).convert_to_fp16()
        enc = gclip.CLIPVisionEncoder().to(dev).eval()
        txt = F.normalize(th.randn((1, 512), device=dev), dim=-1)
        cond_fn, model_fn, ddim = gclip.CLIPGuidance(enc, txt, 100.0), model.eval(), True
    elif args.config == "sr512":
        kw = su.sr_model_and_diffusion_defaults()
        kw.update(large_size=512, small_size=128, num_channels=192, num_res_blocks=2, attention_resolutions="32,16",
                  num_head_channels=64, class_cond=True, learn_sigma=True, resblock_updown=True,
                  use_scale_shift_norm=True, use_fp16=True, timestep_respacing="250")
        model, diffusion = su.sr_create_model_and_diffusion(**kw)
        randomize_(model, 99)
        model.to(dev).convert_to_fp16()
        cond_fn, model_fn = None, model.eval()
        extra = {"low_res": th.rand((B, 3, 128, 128), device=dev) * 2 - 1}
    else:
        kw = unet_kwargs(512)
        kw.update(use_fp16=False, timestep_respacing="ddim25")
        model, diffusion = su.create_model_and_diffusion(**kw)
        randomize_(model, 7)
        model.to(dev).eval()
        ckw = clf_kwargs(512)
        ckw.update(classifier_use_fp16=False)
        clf = su.create_classifier(**ckw)
        randomize_(clf, 8)
        clf.to(dev).eval()
        cond_fn, model_fn, ddim = ClassifierGuidance(clf, 4.0), ModelFn(model, True), True
        dtype = "f16 storage of fp32 masters, fp32 accumulate (use_fp16=False checkpoints run the same kernels)"
    ms, finite, stepper, clocks, _, _, _ = time_steps(diffusion, model_fn, cond_fn, B, S, dev, args.steps, args.warmup,
                                                      world, ddim=ddim, extra_kwargs=extra, clocks_index=local)
    if rank != 0:
        return
    burst, sustained, hbm, src = peaks()
    tf = c["gflop"] * 1e9 * B / (ms * 1e-3) / 1e12
    line = {
        "metric": "samples_per_sec", "value": world * B / (c["steps"] * ms * 1e-3), "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": {"workload": f"BASELINE configs[{c['index']}]: {c['workload']}, batch {B}/GPU", "per_gpu_batch": B,
                   "global_batch": B * world, "steps_per_sample": c["steps"],
                   "l2_policy": "activations per step exceed the 126 MB L2; no flush needed"},
        "finite": finite, "clocks": clocks, "launches_per_step": stepper.launches_per_step,
        "gpu_launches": stepper.launches_per_step * args.steps,
        "roofline": {"bound": "tensor", "kernel": "whole step (conv_igemm_kernel dominates)", "achieved": tf,
                     "peak": sustained, "unit": "TFLOP/s", "frac": tf / sustained, "peak_source": src,
                     "gflop_per_sample_step": c["gflop"], "traffic": None},
    }
    emit(line)


_REAL_STDOUT = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL writes "NCCL version ..." to stdout
    when the box sets NCCL_DEBUG=VERSION): send file descriptor 1 to stderr for the whole run and keep a private
    duplicate of the real stdout for the result line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=["cfg2"] + sorted(OTHER_CONFIGS),
                    help="cfg2 = BASELINE configs[1] (the headline metric); the others are configs[2..4]")
    ap.add_argument("--batch", type=int, default=None,
                    help="per-GPU batch (default 64 = the BASELINE configs[1] batch on one GPU; weak scaling keeps it)")
    ap.add_argument("--strong-global-batch", type=int, default=64)
    ap.add_argument("--no-strong", action="store_true", help="skip the global-batch-64 (64/N per GPU) record")
    ap.add_argument("--no-full-loop", action="store_true", help="skip the whole 250-step loop + gather measurement")
    ap.add_argument("--image-size", type=int, default=256)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.batch_given = args.batch is not None
    if args.batch is None:
        args.batch = 64
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.config != "cfg2":
        run_other_config(args)
    else:
        run_gpu_arm(args)
    if th.distributed.is_available() and th.distributed.is_initialized():
        th.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
