"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_models.py header).

CPU restatement of the reference's diffusion bookkeeping and per-step update:
  space_timesteps            /root/reference/guided_diffusion/respace.py:7-60
  respaced betas / map       respace.py:72-86
  float64 tables             gaussian_diffusion.py:118-169
  p_mean_variance            gaussian_diffusion.py:232-326   (EPSILON|START_X x LEARNED_RANGE|FIXED_*)
  condition_mean / _score    gaussian_diffusion.py:356-393
  p_sample / ddim_sample     gaussian_diffusion.py:395-439, 546-594
The closed forms are those of SURVEY App. B.1.  Integer outputs are pinned bit-exactly and float64 tables to
1e-15 against the real reference by tests/golden/diffusion_golden.json (oracle/make_golden.py).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch as th


def space_timesteps(num_timesteps: int, section_counts) -> List[int]:
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            n = int(section_counts[4:])
            for s in range(1, num_timesteps):
                if len(range(0, num_timesteps, s)) == n:
                    return sorted(range(0, num_timesteps, s))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(v) for v in section_counts.split(",")]
    k = len(section_counts)
    out, start = [], 0
    for i, cnt in enumerate(section_counts):
        size = num_timesteps // k + (1 if i < num_timesteps % k else 0)
        if size < cnt:
            raise ValueError(f"cannot divide section of {size} steps into {cnt}")
        step = 1 if cnt <= 1 else (size - 1) / (cnt - 1)
        cur = 0.0
        for _ in range(cnt):
            out.append(start + round(cur))
            cur += step
        start += size
    return sorted(set(out))


def named_betas(name: str, T: int) -> np.ndarray:
    if name == "linear":
        s = 1000 / T
        return np.linspace(s * 0.0001, s * 0.02, T, dtype=np.float64)
    if name == "cosine":
        f = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
        return np.array([min(1 - f((i + 1) / T) / f(i / T), 0.999) for i in range(T)])
    raise NotImplementedError(name)


class Tables:
    """Respaced process: timestep_map + the float64 tables the sampler reads."""

    def __init__(self, schedule: str = "linear", steps: int = 1000, respacing="", learn_sigma: bool = True,
                 sigma_small: bool = False, predict_xstart: bool = False, rescale_timesteps: bool = False):
        base = named_betas(schedule, steps)
        keep = space_timesteps(steps, respacing if respacing else [steps])
        base_acp = np.cumprod(1.0 - base, axis=0)
        last, betas, tmap = 1.0, [], []
        for i, a in enumerate(base_acp):
            if i in set(keep):
                betas.append(1 - a / last)
                last = a
                tmap.append(i)
        self.timestep_map = tmap
        self.original_num_steps = steps
        self.rescale_timesteps = rescale_timesteps
        b = np.array(betas, dtype=np.float64)
        self.betas = b
        self.T = len(b)
        al = 1.0 - b
        acp = np.cumprod(al, axis=0)
        self.acp, self.acp_prev = acp, np.append(1.0, acp[:-1])
        self.sqrt_recip = np.sqrt(1.0 / acp)
        self.sqrt_recipm1 = np.sqrt(1.0 / acp - 1)
        self.post_var = b * (1.0 - self.acp_prev) / (1.0 - acp)
        self.post_logvar = np.log(np.append(self.post_var[1], self.post_var[1:]))
        self.coef1 = b * np.sqrt(self.acp_prev) / (1.0 - acp)
        self.coef2 = (1.0 - self.acp_prev) * np.sqrt(al) / (1.0 - acp)
        self.learn_sigma, self.sigma_small, self.predict_xstart = learn_sigma, sigma_small, predict_xstart

    def model_t(self, i: int) -> float:
        t = self.timestep_map[i]
        return t * (1000.0 / self.original_num_steps) if self.rescale_timesteps else t

    def _f(self, arr, i) -> th.Tensor:
        return th.tensor(arr[i]).float()  # float64 -> float32, like `.float()` at gaussian_diffusion.py:914

    def mean_variance(self, model_out: th.Tensor, x: th.Tensor, i: int, clip: bool = True,
                      denoised_fn=None) -> Dict[str, th.Tensor]:
        C = x.shape[1]
        if self.learn_sigma:
            eps, v = model_out[:, :C], model_out[:, C:]
            frac = (v + 1) / 2
            logvar = frac * self._f(np.log(self.betas), i) + (1 - frac) * self._f(self.post_logvar, i)
            var = th.exp(logvar)
        else:
            eps = model_out
            if self.sigma_small:
                va, lv = self.post_var, self.post_logvar
            else:
                va = np.append(self.post_var[1], self.betas[1:])
                lv = np.log(va)
            var = self._f(va, i).expand(x.shape)
            logvar = self._f(lv, i).expand(x.shape)
        x0 = eps if self.predict_xstart else self._f(self.sqrt_recip, i) * x - self._f(self.sqrt_recipm1, i) * eps
        if denoised_fn is not None:  # process_xstart, gaussian_diffusion.py:262-265
            x0 = denoised_fn(x0)
        if clip:
            x0 = x0.clamp(-1, 1)
        mean = self._f(self.coef1, i) * x0 + self._f(self.coef2, i) * x
        return {"mean": mean, "variance": var, "log_variance": logvar, "pred_xstart": x0}

    def p_sample(self, model_out, x, i, z, grad: Optional[th.Tensor] = None, clip: bool = True, denoised_fn=None):
        o = self.mean_variance(model_out, x, i, clip, denoised_fn)
        mean = o["mean"]
        if grad is not None:
            mean = mean.float() + o["variance"] * grad.float()
        nz = 0.0 if i == 0 else 1.0
        return {"sample": mean + nz * th.exp(0.5 * o["log_variance"]) * z, "pred_xstart": o["pred_xstart"]}

    def ddim_sample(self, model_out, x, i, z, grad: Optional[th.Tensor] = None, eta: float = 0.0, clip: bool = True,
                    denoised_fn=None):
        o = self.mean_variance(model_out, x, i, clip, denoised_fn)
        sr, srm1 = self._f(self.sqrt_recip, i), self._f(self.sqrt_recipm1, i)
        ab, abp = self._f(self.acp, i), self._f(self.acp_prev, i)
        x0 = o["pred_xstart"]
        if grad is not None:
            e = (sr * x - x0) / srm1
            e = e - (1 - ab).sqrt() * grad
            x0 = sr * x - srm1 * e
        e2 = (sr * x - x0) / srm1
        sigma = eta * th.sqrt((1 - abp) / (1 - ab)) * th.sqrt(1 - ab / abp)
        mean_pred = x0 * th.sqrt(abp) + th.sqrt(1 - abp - sigma ** 2) * e2
        nz = 0.0 if i == 0 else 1.0
        return {"sample": mean_pred + nz * sigma * z, "pred_xstart": x0}


def _ddim_reverse(self, model_out, x, i, clip: bool = True):
    """ddim_reverse_sample, gaussian_diffusion.py:596-632 (eta = 0)."""
    o = self.mean_variance(model_out, x, i, clip)
    sr, srm1 = self._f(self.sqrt_recip, i), self._f(self.sqrt_recipm1, i)
    eps = (sr * x - o["pred_xstart"]) / srm1
    nxt = self._f(np.append(self.acp[1:], 0.0), i)
    return {"sample": o["pred_xstart"] * th.sqrt(nxt) + th.sqrt(1 - nxt) * eps, "pred_xstart": o["pred_xstart"]}


Tables.ddim_reverse_sample = _ddim_reverse


def to_uint8_nhwc(sample: th.Tensor) -> th.Tensor:
    """scripts/classifier_sample.py:87-89."""
    return ((sample + 1) * 127.5).clamp(0, 255).to(th.uint8).permute(0, 2, 3, 1).contiguous()


def driver_order(num_samples: int, batch_size: int, world: int):
    """Global sample index -> (iteration, rank, index in batch) of scripts/classifier_sample.py:70-102."""
    out, it = [], 0
    lists = 0
    while lists * batch_size < num_samples:
        for r in range(world):
            for j in range(batch_size):
                out.append((it, r, j))
        lists += world
        it += 1
    return out[:num_samples], it
