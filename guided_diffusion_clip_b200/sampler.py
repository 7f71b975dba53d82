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
        """Stepper for this (diffusion, model, guidance, shape) or None when the fast path does not apply.
        Cached on the diffusion object; invalidated when either model's parameters change."""
        if os.environ.get("GD_B200_NO_GRAPH", "0") == "1" or th.device(device).type != "cuda":
            return None
        m = model.model if isinstance(model, ModelFn) else model
        if type(m) is not UNetModel:
            return None
        if cond_fn is not None and not isinstance(cond_fn, ClassifierGuidance):
            return None
        key = (id(m), m._param_version, getattr(model, "class_cond", True),
               (id(cond_fn.classifier), cond_fn.classifier._param_version, cond_fn.classifier_scale) if cond_fn else None,
               tuple(shape), str(device), bool(clip_denoised), bool(ddim), float(eta))
        cache = diffusion.__dict__.setdefault("_steppers", {})
        if key not in cache:
            cache[key] = GraphedStepper.maybe_create(diffusion, model, cond_fn, shape, device, model_kwargs,
                                                     clip_denoised, ddim, eta)
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
        from .engine import UNetPlan, norm_device
        device = norm_device(device)
        self.unet = model._plan_for(("unet", n, h, w, str(device)), lambda: UNetPlan(model, n, h, w, device))
        self.clf = None
        self.scale = 0.0
        if cond_fn is not None:
            self.clf = cond_fn.classifier.plan(n, h, w, th.device(device))
            self.scale = cond_fn.classifier_scale
        dev = th.device(device)
        self.t_idx = th.zeros((n,), dtype=th.int64, device=dev)
        self.noise = th.empty(shape, dtype=th.float32, device=dev)
        self.sample = th.empty(shape, dtype=th.float32, device=dev)
        self.x0 = th.empty(shape, dtype=th.float32, device=dev)
        self.y = y.to(device=dev, dtype=th.int64).contiguous().clone() if isinstance(y, th.Tensor) else None
        self.uses_y = uses_y
        if uses_y:
            self.unet.cond_in.copy_(self.y)
        self.map = diffusion.map_tensor(dev) if hasattr(diffusion, "map_tensor") else None
        self.rescale = (1000.0 / diffusion.original_num_steps if (self.map is not None and diffusion.rescale_timesteps)
                        else (1000.0 / diffusion.num_timesteps if diffusion.rescale_timesteps else None))
        self.clip, self.ddim, self.eta = clip_denoised, ddim, eta
        self.graph: Optional[th.cuda.CUDAGraph] = None
        self.overlap = os.environ.get("GD_B200_NO_OVERLAP", "0") != "1"
        self._fork = th.cuda.Stream(device=dev)
        self._capture()

    def _body(self) -> None:
        d = self.diffusion
        ts = self.map[self.t_idx] if self.map is not None else self.t_idx
        ts = ts.float() * self.rescale if self.rescale is not None else ts
        self.unet.t_in.copy_(ts)
        grad = None
        if self.clf is not None and self.overlap:
            # The classifier (fwd + bwd) and the UNet both depend only on (x_t, t): fork the classifier onto a
            # second stream so its bandwidth-bound GroupNorm passes fill in beside the UNet's tensor-bound convs
            # (and vice versa); join before the posterior kernel.  Captured as two branches of the step graph.
            main = th.cuda.current_stream()
            self._fork.wait_stream(main)
            with th.cuda.stream(self._fork):
                self.clf.x_in.copy_(self.unet.x_in)
                self.clf.t_in.copy_(self.unet.t_in)
                grad = self.clf.guidance(self.clf.x_in, self.clf.t_in, self.y, self.scale)
            self.unet.prog.run()
            main.wait_stream(self._fork)
        else:
            self.unet.prog.run()
            if self.clf is not None:
                self.clf.x_in.copy_(self.unet.x_in)
                self.clf.t_in.copy_(ts)
                grad = self.clf.guidance(self.clf.x_in, self.clf.t_in, self.y, self.scale)
        d._launch_posterior(x=self.unet.x_in, t=self.t_idx, model_out=self.unet.out, grad=grad, noise=self.noise,
                            sample=self.sample, pred_xstart=self.x0, clip_denoised=self.clip, ddim=self.ddim,
                            eta=self.eta)

    def _capture(self) -> None:
        # warm-up on a side stream (sets function attributes, resolves driver entry points), then capture
        self.unet.x_in.zero_()
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
        n = self.unet.prog.launches + 1
        if self.clf is not None:
            n += self.clf.fwd.launches + self.clf.bwd.launches + 1
        return n

    def step(self, img: th.Tensor, t: th.Tensor, noise: Optional[th.Tensor] = None,
             labels: Optional[th.Tensor] = None):
        if labels is not None and self.y is not None:
            self.y.copy_(labels)
            if self.uses_y:
                self.unet.cond_in.copy_(self.y)
        self.unet.x_in.copy_(img)
        self.t_idx.copy_(t)
        if noise is None:
            self.noise.normal_()
        else:
            self.noise.copy_(noise)
        self.graph.replay()
        return {"sample": self.sample.clone(), "pred_xstart": self.x0.clone()}
