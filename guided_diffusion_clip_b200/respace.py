"""Timestep respacing (guided_diffusion/respace.py:7-128): which of the original T steps are kept, the betas
of the shortened chain, and the index -> original-timestep map applied to BOTH the model and cond_fn.
All of this is integer / float64 bookkeeping on the host and must match the reference bit for bit."""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch as th

from .gaussian_diffusion import GaussianDiffusion


def space_timesteps(num_timesteps, section_counts):
    """Return the set of kept timesteps (respace.py:7-60).

    "ddimN" -> the fixed-stride subset of the DDIM paper (exactly N steps or ValueError);
    otherwise a comma list / list of per-section counts, each section strided with a fractional
    stride and Python's round() (banker's rounding — part of the contract, respace.py:56)."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                if len(range(0, num_timesteps, stride)) == want:
                    return set(range(0, num_timesteps, stride))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    n_sec = len(section_counts)
    base, extra = divmod(num_timesteps, n_sec)
    kept: List[int] = []
    start = 0
    for i, count in enumerate(section_counts):
        size = base + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        pos = 0.0
        for _ in range(count):
            kept.append(start + round(pos))
            pos += stride
        start += size
    return set(kept)


class SpacedDiffusion(GaussianDiffusion):
    """A diffusion process over a subset of the base process's timesteps (respace.py:63-113)."""

    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.timestep_map: List[int] = []
        self.original_num_steps = len(kwargs["betas"])
        base = GaussianDiffusion(**kwargs)
        prev = 1.0
        new_betas = []
        for i, acp in enumerate(base.alphas_cumprod):
            if i in self.use_timesteps:
                new_betas.append(1 - acp / prev)
                prev = acp
                self.timestep_map.append(i)
        kwargs["betas"] = np.array(new_betas)
        super().__init__(**kwargs)
        self._map_cache: Dict[str, th.Tensor] = {}

    def map_tensor(self, device) -> th.Tensor:
        """int64 device copy of timestep_map, uploaded once per device (the reference re-uploads per call)."""
        key = str(device)
        t = self._map_cache.get(key)
        if t is None:
            t = th.tensor(self.timestep_map, device=device, dtype=th.int64)
            self._map_cache[key] = t
        return t

    def _wrap(self, fn):
        if isinstance(fn, _WrappedModel):
            return fn
        return _WrappedModel(fn, self.timestep_map, self.rescale_timesteps, self.original_num_steps, self)

    _wrap_model = _wrap  # reference name (respace.py:104)

    def _scale_timesteps(self, t):
        return t  # scaling happens inside the wrapper (respace.py:111-113)


class _WrappedModel:
    """Calls `model(x, timestep_map[ts])` (respace.py:116-128)."""

    def __init__(self, model, timestep_map, rescale_timesteps, original_num_steps, owner=None):
        self.model = model
        self.timestep_map = timestep_map
        self.rescale_timesteps = rescale_timesteps
        self.original_num_steps = original_num_steps
        self._owner = owner

    def __call__(self, x, ts, **kwargs):
        if self._owner is not None:
            map_tensor = self._owner.map_tensor(ts.device).to(ts.dtype)
        else:
            map_tensor = th.tensor(self.timestep_map, device=ts.device, dtype=ts.dtype)
        new_ts = map_tensor[ts]
        if self.rescale_timesteps:
            new_ts = new_ts.float() * (1000.0 / self.original_num_steps)
        return self.model(x, new_ts, **kwargs)
