#!/usr/bin/env python
"""bench.py — 256x256 classifier-guided ADM sampling (BASELINE.json configs[1]) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU cores

A "step" is one guided sampling step on one per-GPU batch: UNet-256 forward + classifier-256 forward + classifier
data-gradient backward + fused posterior/noise update (gaussian_diffusion.py:395-439 of the reference).  A sample
needs 250 such steps (timestep_respacing="250"), so  value [samples/s] = n_gpus * batch / (250 * s_per_step).
Weak scaling: the per-GPU batch is fixed at 64 (N=1 is exactly the BASELINE configs[1] batch; --batch 8 gives the
"global 64 on 8 GPUs" split); ranks are independent (no data-path collective); the one collective of the path, the
all_gather of finished uint8 samples, is exercised after the timed region and reported as gather_ms.
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


def workload_config(image_size, batch, world):
    """The `config` object shared by both arms (the reference arm times a bounded sample of the same workload)."""
    return {"workload": f"{image_size}x{image_size} class-cond ADM (256ch, 2 res blocks, attn 32/16/8) + EncoderUNet "
                        f"classifier guidance (scale 1.0), 250 respaced steps, batch {batch}/GPU (global {batch * world}), "
                        "fp16 storage fp32 accumulate",
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
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port) on the host cores, bounded sample
# ------------------------------------------------------------------------------------------------------------
def cpu_guided_steps(n_steps: int, image_size: int, batch: int = 1):
    """Time `n_steps` guided steps of the reference algorithm (oracle/, a CPU restatement pinned to the reference by
    tests/golden) at batch `batch` with all host threads.  Returns seconds per step."""
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


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times = cpu_guided_steps(args.warmup + args.steps, args.image_size, 1)
    timed = times[args.warmup:]
    s_per_step = sum(timed) / len(timed)
    value = 1.0 / (STEPS_PER_SAMPLE * s_per_step)
    cores = os.cpu_count() or 1
    sample = (f"{len(timed)} guided steps at batch 1, {args.image_size}x{args.image_size}, fp32 oneDNN, "
              f"{cores} threads; samples/s extrapolated x{STEPS_PER_SAMPLE} steps")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {
        "impl": "reference", "metric": "guided_samples_per_sec_256", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.image_size, args.batch, world),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
def top_conv_roofline(batch, image_size, burst_tflops, reps=12):
    """The dominant kernel: 3x3 conv 256->256 at full resolution (31 % of the step's FLOPs), timed
    alone with CUDA events on the launching stream.  Inputs (batch*H*W*256 fp16 = 268 MB at batch 8) exceed the
    126 MB L2, and launches rotate over two input buffers."""
    import ctypes as C
    from guided_diffusion_clip_b200 import _lib as L
    from guided_diffusion_clip_b200.engine import pack_conv3x3
    lib = L.load()
    c = 256
    xs = [th.randn((batch, image_size, image_size, c), device="cuda", dtype=th.float16) for _ in range(2)]
    out = th.empty_like(xs[0])
    w = pack_conv3x3(th.randn((c, c, 3, 3), device="cuda") * 0.02)
    bias = th.zeros(c, device="cuda")
    d = L.ConvDesc()
    d.c0, d.ld0, d.taps, d.n, d.h, d.w = c, c, 9, batch, image_size, image_size
    d.wpack, d.k_total, d.n_pad, d.bias, d.cout = w.data_ptr(), 9 * c, c, bias.data_ptr(), c
    d.out, d.ld_out, d.out_mode, d.out_scale = out.data_ptr(), c, L.OUT_NHWC_F16, 1.0
    stream = C.c_void_p(th.cuda.current_stream().cuda_stream)
    evs = []
    for i in range(reps + 3):
        d.a0 = xs[i % 2].data_ptr()
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.gd_conv_igemm(C.byref(d), stream), "gd_conv_igemm")
        e1.record()
        evs.append((e0, e1))
    th.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in evs[3:]]
    avg = sum(ms) / len(ms)
    flops = 2.0 * batch * image_size * image_size * c * 9 * c
    achieved = flops / (avg * 1e-3) / 1e12
    # DRAM traffic of this launch from the committed ncu capture: 2.209 GB read + 2.173 GB written at batch 64, i.e.
    # 1.02x the algorithmic bytes (activations in + out once, weights once)
    algo_bytes = 2.0 * batch * image_size * image_size * c * 2 + 9 * c * c * 2
    traffic = (2.209482e9 + 2.173346e9) * batch / 64.0 if image_size == 256 else None
    return {"bound": "tensor", "kernel": "conv_igemm_kernel 3x3 256->256 @%dx%d batch %d" % (image_size, image_size, batch),
            "achieved": achieved, "peak": burst_tflops, "unit": "TFLOP/s", "frac": achieved / burst_tflops,
            "traffic": traffic, "traffic_unit": "bytes/launch", "traffic_source": "dram__bytes_read.sum+dram__bytes_write.sum, "
            "ncu --set full, profiles/ncu_conv_r01i_halo_batch64.txt (batch 64; scaled by batch)",
            "algorithmic_bytes": algo_bytes, "tensor_pipe_active_pct_ncu": 99.56,
            "avg_launch_ms": avg, "flops_per_launch": flops}


def hbm_kernels_roofline(diffusion, batch, image_size, hbm_gbs, reps=10):
    """The two bandwidth-bound kernels of the step timed alone (CUDA events, current stream), rotating over 3 input
    sets so that no launch finds its inputs in the 126 MB L2: the fused posterior update (21 channels x 4 B per
    pixel = reads x, eps|v, grad, noise; writes sample, pred_xstart) and GroupNorm+SiLU+FiLM apply on the largest
    activation (256 channels at full resolution: read fp16 once, write fp16 once)."""
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
    out.append({"kernel": "posterior_kernel (guided p_sample update)", "bound": "hbm", "achieved": gbs, "peak": hbm_gbs,
                "unit": "GB/s", "frac": gbs / hbm_gbs, "avg_launch_ms": ms, "algorithmic_bytes": nbytes})
    del sets
    c = 256
    xs = [th.randn((batch, image_size, image_size, c), device=dev, dtype=th.float16) for _ in range(3)]
    y = th.empty_like(xs[0])
    st = th.zeros((batch, 32, 2), device=dev)
    st[..., 1] = 1.0
    gamma, beta = th.ones(c, device=dev), th.zeros(c, device=dev)
    film = th.randn((batch, 2 * c), device=dev) * 0.1
    stream = C.c_void_p(th.cuda.current_stream().cuda_stream)
    vp = lambda v: C.c_void_p(v.data_ptr())  # noqa: E731

    def gn(i):
        L.check(lib.gd_groupnorm_apply(vp(xs[i % 3]), c, vp(st), vp(gamma), vp(beta), vp(film), 2 * c, vp(y), c, batch,
                                       image_size, image_size, c, 1, L.GN_SAME, None, 0, stream), "gd_groupnorm_apply")

    ms = timed(gn)
    nbytes = 2.0 * 2 * batch * image_size * image_size * c
    gbs = nbytes / (ms * 1e-3) / 1e9
    out.append({"kernel": "gn_apply_kernel (GroupNorm32+FiLM+SiLU, 256 ch @%dx%d)" % (image_size, image_size),
                "bound": "hbm", "achieved": gbs, "peak": hbm_gbs, "unit": "GB/s", "frac": gbs / hbm_gbs,
                "avg_launch_ms": ms, "algorithmic_bytes": nbytes})
    return out


def run_gpu_arm(args):
    from guided_diffusion_clip_b200 import dist_util, script_util as su
    from guided_diffusion_clip_b200.sampler import ClassifierGuidance, GraphedStepper, ModelFn
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
    th.manual_seed(dist_util.rank_seed(args.seed, rank))

    ukw = unet_kwargs(S)
    model, diffusion = su.create_model_and_diffusion(**ukw)
    randomize_(model, 1234)
    model.to(dev)
    model.convert_to_fp16()
    model.eval()
    classifier = su.create_classifier(**clf_kwargs(S))
    randomize_(classifier, 4321)
    classifier.to(dev)
    classifier.convert_to_fp16()
    classifier.eval()
    cond_fn = ClassifierGuidance(classifier, 1.0)
    model_fn = ModelFn(model, True)
    shape = (B, 3, S, S)
    y = th.randint(0, 1000, (B,), device=dev)
    # the public per-step API (diffusion.p_sample) resolves to a cached CUDA-graph replay for our own objects
    stepper = GraphedStepper.cached(diffusion, model_fn, cond_fn, shape, dev, {"y": y}, True, False, 0.0)
    assert stepper is not None, "fast path not taken"
    launches_per_step = stepper.launches_per_step

    T = diffusion.num_timesteps
    img = th.randn(*shape, device=dev)
    t = th.empty((B,), dtype=th.int64, device=dev)
    kwargs = {"y": y}

    def one_step(k):
        nonlocal img
        t.fill_(T - 1 - (k % T))
        img = diffusion.p_sample(model_fn, img, t, cond_fn=cond_fn, model_kwargs=kwargs)["sample"]

    for k in range(args.warmup):
        one_step(k)
    th.cuda.synchronize()
    if world > 1:
        dist.barrier()
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record()
        for k in range(args.steps):
            one_step(args.warmup + k)
        e1.record()
        th.cuda.synchronize()
    if world > 1:
        dist.barrier()
    th.cuda.synchronize()
    ms_total = th.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms_total) / args.steps
    finite = bool(th.isfinite(img).all())

    # ---- end-to-end through the public API with HOST buffers: H2D of the step's inputs, D2H of its result -------
    host_x = th.randn(*shape).pin_memory()
    host_t = th.empty((B,), dtype=th.int64).pin_memory()
    host_out = th.empty(shape).pin_memory()
    dx = th.empty(shape, device=dev)
    dt = th.empty((B,), dtype=th.int64, device=dev)

    def e2e_step(k):
        host_t.fill_(T - 1 - (k % T))
        dx.copy_(host_x, non_blocking=True)
        dt.copy_(host_t, non_blocking=True)
        out = diffusion.p_sample(model_fn, dx, dt, cond_fn=cond_fn, model_kwargs=kwargs)
        host_out.copy_(out["sample"], non_blocking=True)
        th.cuda.current_stream().synchronize()
        host_x.copy_(host_out)

    for k in range(min(args.warmup, 3)):
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

    # ---- the path's one collective: all_gather of the finished uint8 batch (outside the timed steps) -----------
    g0, g1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    u8 = dist_util.to_uint8_nhwc(img)
    g0.record()
    imgs, labs = dist_util.all_gather_batch(u8, y)
    g1.record()
    th.cuda.synchronize()
    gather_ms = g0.elapsed_time(g1)

    if rank != 0:
        return
    burst, sustained, hbm, src = peaks()
    roof = top_conv_roofline(B, S, burst)
    roof["peak_source"] = src
    step_tflops = GFLOP_PER_SAMPLE_STEP * 1e9 * B / (ms_per_step * 1e-3) / 1e12
    roof["step_tflops"] = step_tflops
    roof["step_frac_of_sustained"] = step_tflops / sustained
    roof_hbm = hbm_kernels_roofline(diffusion, B, S, hbm)
    value = world * B / (STEPS_PER_SAMPLE * ms_per_step * 1e-3)
    line = {
        "metric": "guided_samples_per_sec_256", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": workload_config(S, B, world), "finite": finite,
        "roofline": roof, "roofline_hbm": roof_hbm, "clocks": clocks.summary(), "gpu_launches": launches_per_step * args.steps,
        "launches_per_step": launches_per_step, "gather_ms": gather_ms,
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": bytes_io + B * 8,
                "d2h_bytes_per_step": bytes_io, "steps": n_e2e},
    }
    if world == 1 and not args.no_cpu_baseline:
        times = cpu_guided_steps(1 + args.cpu_steps, S, 1)
        s = sum(times[1:]) / len(times[1:])
        cores = os.cpu_count() or 1
        line["cpu_baseline"] = {
            "value": 1.0 / (STEPS_PER_SAMPLE * s), "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": f"{args.cpu_steps} guided step(s) at batch 1 after 1 warm-up, {S}x{S}, fp32, {cores} threads; "
                      f"extrapolated x{STEPS_PER_SAMPLE} steps ({s:.2f} s/step)"}
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
    ap.add_argument("--batch", type=int, default=64,
                    help="per-GPU batch (64 = the BASELINE configs[1] batch on one GPU; weak scaling keeps it per GPU)")
    ap.add_argument("--image-size", type=int, default=256)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-generic", action="store_true", help="e2e through diffusion.p_sample instead of the stepper")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)
    if th.distributed.is_available() and th.distributed.is_initialized():
        th.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
