"""Guidance objects and the CUDA-graph fast path of the sampling loop.

* ClassifierGuidance is the `cond_fn` of scripts/classifier_sample.py:54-61 as an object: it evaluates
  scale * d/dx sum_b log_softmax(classifier(x, t))[b, y_b] with the classifier plan's forward + dX programs.
* GraphedStepper captures one whole guided step (UNet forward -> classifier forward -> classifier dX backward
  -> fused posterior update) into a CUDA graph when both the model and the cond_fn are ours; per step the
  host then does three small copies, one RNG draw and one graph launch instead of ~600 launches.
  Arbitrary Python callables (e.g. the reference's own model_fn / cond_fn closures) still work through the
  generic path in gaussian_diffusion.GaussianDiffusion._sample_step.
"""
from __future__ import annotations

import os
from typing import Optional

import torch as th

from . import _lib as L
from .unet import EncoderUNetModel, UNetModel


class ClassifierGuidance:
    """cond_fn(x, t, y=...) -> classifier_scale * grad_x log p(y | x_t, t)."""

    def __init__(self, classifier: EncoderUNetModel, classifier_scale: float = 1.0):
        self.classifier = classifier
        self.classifier_scale = float(classifier_scale)

    def __call__(self, x, t, y=None, **kwargs):
        assert y is not None
        n, _, h, w = x.shape
        plan = self.classifier.plan(n, h, w, x.device)
        return plan.guidance(x, t, y, self.classifier_scale).clone()


class ModelFn:
    """model_fn of scripts/classifier_sample.py:63-65: forwards `y` only when the model is class-conditional."""

    def __init__(self, model: UNetModel, class_cond: bool = True):
        self.model = model
        self.class_cond = class_cond

    def __call__(self, x, t, y=None, **kwargs):
        return self.model(x, t, y if self.class_cond else None)

    def parameters(self):
        return self.model.parameters()


class GraphedStepper:
    """One sampling step as a single CUDA-graph replay (only for our own model / guidance objects)."""

    @staticmethod
    def cached(diffusion, model, cond_fn, shape, device, model_kwargs, clip_denoised, ddim, eta):
        """Stepper for this (diffusion, model, guidance, shape) or None when no fast path applies.  Cached on the
        diffusion object; invalidated when any involved model's parameters change.  Two kinds: the specialised
        GraphedStepper (class-conditional / unconditional UNetModel + ClassifierGuidance: classifier forked onto a
        second stream) and GenericGraphedStepper (every other combination of OUR model classes and guidance objects,
        e.g. SuperResModel + low_res, the fork's clip_feat models, CLIPGuidance)."""
        if os.environ.get("GD_B200_NO_GRAPH", "0") == "1" or th.device(device).type != "cuda":
            return None
        m = model.model if isinstance(model, ModelFn) else model
        if not isinstance(m, UNetModel):
            return None
        sig = GenericGraphedStepper.signature(m, model, cond_fn, model_kwargs, device)
        if sig is None:
            return None
        key = (sig, tuple(shape), str(device), bool(clip_denoised), bool(ddim), float(eta))
        cache = diffusion.__dict__.setdefault("_steppers", {})
        if key not in cache:
            # steppers captured for an older parameter version of the same model object can never be hit again: drop
            # them (each one owns a CUDA graph and its static buffers)
            for stale in [k for k in cache if k[0][0] == sig[0] and k[0][1] != sig[1]]:
                del cache[stale]
            st = GraphedStepper.maybe_create(diffusion, model, cond_fn, shape, device, model_kwargs, clip_denoised,
                                             ddim, eta)
            if st is None:
                st = GenericGraphedStepper(diffusion, model, cond_fn, tuple(shape), device, model_kwargs,
                                           clip_denoised, ddim, eta)
            cache[key] = st
        return cache[key]

    @staticmethod
    def maybe_create(diffusion, model, cond_fn, shape, device, model_kwargs, clip_denoised, ddim, eta):
        if os.environ.get("GD_B200_NO_GRAPH", "0") == "1":
            return None
        if th.device(device).type != "cuda":
            return None
        class_cond = True
        if isinstance(model, ModelFn):
            class_cond = model.class_cond
            model = model.model
        if type(model) is not UNetModel or model.label_mlp:
            return None
        if cond_fn is not None and not isinstance(cond_fn, ClassifierGuidance):
            return None
        kw = dict(model_kwargs or {})
        y = kw.pop("y", None)
        if kw:
            return None
        uses_y = model.num_classes is not None and class_cond
        if (uses_y or cond_fn is not None) and not isinstance(y, th.Tensor):
            return None
        if (model.num_classes is not None) != uses_y:
            return None
        return GraphedStepper(diffusion, model, cond_fn, tuple(shape), device, y, uses_y, clip_denoised, ddim, eta)

    def __init__(self, diffusion, model, cond_fn, shape, device, y, uses_y, clip_denoised, ddim, eta):
        n, c, h, w = shape
        self.diffusion = diffusion
        from .engine import ClassifierPlan, UNetPlan, norm_device
        device = norm_device(device)
        dev = th.device(device)
        # Samples are independent, so the batch can be cut into `split` contiguous parts that run as independent
        # branches of the step graph: one part's bandwidth-bound GroupNorm passes then overlap another part's
        # tensor-bound convolutions (GD_B200_SPLIT; 1 = one part).
        split = max(1, int(os.environ.get("GD_B200_SPLIT", "1")))
        if n % split != 0:
            split = 1
        m = n // split
        self.scale = cond_fn.classifier_scale if cond_fn is not None else 0.0
        self.parts = []
        for i in range(split):
            tag = () if i == 0 else (i,)
            unet = model._plan_for(("unet", m, h, w, str(device)) + tag, lambda: UNetPlan(model, m, h, w, device))
            clf = None
            if cond_fn is not None:
                cm = cond_fn.classifier
                clf = cm._plan_for(("clf", m, h, w, str(device)) + tag, lambda: ClassifierPlan(cm, m, h, w, device))
            self.parts.append((slice(i * m, (i + 1) * m), unet, clf, th.cuda.Stream(device=dev), th.cuda.Stream(device=dev)))
        self.unet, self.clf = self.parts[0][1], self.parts[0][2]
        self.t_idx = th.zeros((n,), dtype=th.int64, device=dev)
        self.noise = th.empty(shape, dtype=th.float32, device=dev)
        self.sample = th.empty(shape, dtype=th.float32, device=dev)
        self.x0 = th.empty(shape, dtype=th.float32, device=dev)
        self.y = y.to(device=dev, dtype=th.int64).contiguous().clone() if isinstance(y, th.Tensor) else None
        self.uses_y = uses_y
        if uses_y:
            for sl, unet, _, _, _ in self.parts:
                unet.cond_in.copy_(self.y[sl])
        self.map = diffusion.map_tensor(dev) if hasattr(diffusion, "map_tensor") else None
        self.rescale = (1000.0 / diffusion.original_num_steps if (self.map is not None and diffusion.rescale_timesteps)
                        else (1000.0 / diffusion.num_timesteps if diffusion.rescale_timesteps else None))
        self.clip, self.ddim, self.eta = clip_denoised, ddim, eta
        self.graph: Optional[th.cuda.CUDAGraph] = None
        self.overlap = os.environ.get("GD_B200_NO_OVERLAP", "0") != "1"
        self._capture()

    def _part(self, ts, sl, unet, clf, fork) -> None:
        """One part of the batch on the current stream (+ its classifier on `fork`)."""
        d = self.diffusion
        unet.t_in.copy_(ts[sl])
        y = self.y[sl] if self.y is not None else None
        grad = None
        if clf is not None and self.overlap:
            # The classifier (fwd + bwd) and the UNet both depend only on (x_t, t): fork the classifier onto a
            # second stream so its bandwidth-bound GroupNorm passes fill in beside the UNet's tensor-bound convs
            # (and vice versa); join before the posterior kernel.  Captured as two branches of the step graph.
            cur = th.cuda.current_stream()
            fork.wait_stream(cur)
            with th.cuda.stream(fork):
                clf.x_in.copy_(unet.x_in)
                clf.t_in.copy_(unet.t_in)
                grad = clf.guidance(clf.x_in, clf.t_in, y, self.scale)
            unet.prog.run()
            cur.wait_stream(fork)
        else:
            unet.prog.run()
            if clf is not None:
                clf.x_in.copy_(unet.x_in)
                clf.t_in.copy_(unet.t_in)
                grad = clf.guidance(clf.x_in, clf.t_in, y, self.scale)
        d._launch_posterior(x=unet.x_in, t=self.t_idx[sl], model_out=unet.out, grad=grad, noise=self.noise[sl],
                            sample=self.sample[sl], pred_xstart=self.x0[sl], clip_denoised=self.clip, ddim=self.ddim,
                            eta=self.eta)

    def _body(self) -> None:
        ts = self.map[self.t_idx] if self.map is not None else self.t_idx
        ts = ts.float() * self.rescale if self.rescale is not None else ts
        main = th.cuda.current_stream()
        for _, _, _, side, _ in self.parts[1:]:
            side.wait_stream(main)
        for i, (sl, unet, clf, side, fork) in enumerate(self.parts):
            with th.cuda.stream(main if i == 0 else side):
                self._part(ts, sl, unet, clf, fork)
        for _, _, _, side, _ in self.parts[1:]:
            main.wait_stream(side)

    def _capture(self) -> None:
        # warm-up on a side stream (sets function attributes, resolves driver entry points), then capture
        for _, unet, _, _, _ in self.parts:
            unet.x_in.zero_()
        self.noise.zero_()
        side = th.cuda.Stream()
        side.wait_stream(th.cuda.current_stream())
        with th.cuda.stream(side):
            self._body()
        th.cuda.current_stream().wait_stream(side)
        th.cuda.synchronize()
        self.graph = th.cuda.CUDAGraph()
        with th.cuda.graph(self.graph):
            self._body()

    @property
    def launches_per_step(self) -> int:
        n = 0
        for _, unet, clf, _, _ in self.parts:
            n += unet.prog.launches + 1
            if clf is not None:
                n += clf.fwd.launches + clf.bwd.launches + 1
        return n

    def step(self, img: th.Tensor, t: th.Tensor, noise: Optional[th.Tensor] = None,
             labels: Optional[th.Tensor] = None, model_kwargs=None):
        if labels is not None and self.y is not None:
            self.y.copy_(labels)
            if self.uses_y:
                for sl, unet, _, _, _ in self.parts:
                    unet.cond_in.copy_(self.y[sl])
        for sl, unet, _, _, _ in self.parts:
            unet.x_in.copy_(img[sl])
        self.t_idx.copy_(t)
        if noise is None:
            self.noise.normal_()
        else:
            self.noise.copy_(noise)
        self.graph.replay()
        return {"sample": self.sample.clone(), "pred_xstart": self.x0.clone()}


class GenericGraphedStepper:
    """One sampling step of ANY of our own model classes (UNetModel, SuperResModel, the fork's UNetModel_clip_feat /
    SRImageModel_Feat, optionally behind ModelFn) with no guidance, ClassifierGuidance or CLIPGuidance, as a single
    CUDA-graph replay.  The generic per-step code of GaussianDiffusion._sample_step — model call through the respacing
    wrapper, cond_fn call, fused update — is captured once over static copies of x, t, the noise and every tensor in
    model_kwargs; per step the host copies the inputs into those buffers, draws the noise and replays.  The noise is
    drawn with the same generator call count as the eager path (one normal_ of x's shape per step), so eager and
    graphed loops produce identical bits."""

    @staticmethod
    def signature(m, model, cond_fn, model_kwargs, device):
        """Hashable description of everything the captured graph depends on, or None if something is not ours."""
        from .clip import CLIPGuidance
        dev = th.device(device)
        if cond_fn is None:
            csig = None
        elif isinstance(cond_fn, ClassifierGuidance):
            csig = ("clf", id(cond_fn.classifier), cond_fn.classifier._param_version, cond_fn.classifier_scale)
        elif isinstance(cond_fn, CLIPGuidance):
            if not (isinstance(cond_fn.text, th.Tensor) and cond_fn.text.device.type == "cuda"):
                return None
            csig = ("clip", id(cond_fn.encoder), getattr(cond_fn.encoder, "_param_version", 0), id(cond_fn.text),
                    cond_fn.scale)
        else:
            return None
        ksig = []
        for k in sorted(model_kwargs or {}):
            v = model_kwargs[k]
            if isinstance(v, th.Tensor):
                if v.device.type != dev.type:
                    return None
                ksig.append((k, tuple(v.shape), str(v.dtype)))
            elif v is None or isinstance(v, (bool, int, float, str)):
                ksig.append((k, v))
            else:
                return None
        return (id(m), m._param_version, type(model).__name__, getattr(model, "class_cond", None), csig, tuple(ksig))

    def __init__(self, diffusion, model, cond_fn, shape, device, model_kwargs, clip_denoised, ddim, eta):
        from .engine import norm_device
        dev = th.device(norm_device(device))
        self.diffusion, self.model, self.cond_fn = diffusion, model, cond_fn
        self.clip, self.ddim, self.eta = clip_denoised, ddim, eta
        self.x = th.zeros(shape, dtype=th.float32, device=dev)
        self.t = th.zeros((shape[0],), dtype=th.int64, device=dev)
        self.noise = th.zeros(shape, dtype=th.float32, device=dev)
        self.sample = th.empty(shape, dtype=th.float32, device=dev)
        self.x0 = th.empty(shape, dtype=th.float32, device=dev)
        self.kw = {k: (v.detach().clone() if isinstance(v, th.Tensor) else v) for k, v in (model_kwargs or {}).items()}
        self.launches_per_step = 0
        # eager warm-up on a side stream: builds every plan (host-side weight packing, allocations) outside the capture
        side = th.cuda.Stream(device=dev)
        side.wait_stream(th.cuda.current_stream())
        with th.cuda.stream(side):
            lib = L.load()
            lib.gd_launch_count_reset()
            self._body()
            self.launches_per_step = int(lib.gd_launch_count())
        th.cuda.current_stream().wait_stream(side)
        th.cuda.synchronize()
        self.graph = th.cuda.CUDAGraph()
        with th.cuda.graph(self.graph):
            self._body()

    def _body(self) -> None:
        d = self.diffusion
        with th.no_grad():
            model_out = d._call_model(self.model, self.x, self.t, self.kw)
            grad = None
            if self.cond_fn is not None:
                grad = d._wrap(self.cond_fn)(self.x, d._scale_timesteps(self.t), **self.kw).float().contiguous()
            d._launch_posterior(x=self.x, t=self.t, model_out=model_out, grad=grad, noise=self.noise, sample=self.sample,
                                pred_xstart=self.x0, clip_denoised=self.clip, ddim=self.ddim, eta=self.eta)

    def step(self, img: th.Tensor, t: th.Tensor, noise: Optional[th.Tensor] = None, labels=None, model_kwargs=None):
        self.x.copy_(img)
        self.t.copy_(t)
        for k, v in (model_kwargs or {}).items():
            if isinstance(v, th.Tensor):
                self.kw[k].copy_(v)
        if noise is None:
            self.noise.normal_()
        else:
            self.noise.copy_(noise)
        self.graph.replay()
        return {"sample": self.sample.clone(), "pred_xstart": self.x0.clone()}
