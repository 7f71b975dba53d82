"""Names the reference exposes from guided_diffusion/nn.py that callers of the sampling path touch.
The arithmetic lives in CUDA (csrc/); these are thin host-side entry points."""
from __future__ import annotations

import ctypes as C

import torch as th

from . import _lib as L


def timestep_embedding(timesteps, dim, max_period=10000):
    """Sinusoidal embedding [N, dim] = [cos | sin] (nn.py:103-121) computed by gd_timestep_embedding."""
    if max_period != 10000:
        raise NotImplementedError("only max_period=10000 is compiled in")
    if timesteps.device.type != "cuda":
        raise L.GdError("timestep_embedding only runs on CUDA; there is no CPU path")
    t = timesteps.float().contiguous()
    out = th.empty((t.shape[0], dim), dtype=th.float32, device=t.device)
    with th.cuda.device(t.device):
        stream = C.c_void_p(th.cuda.current_stream().cuda_stream)
        L.check(L.load().gd_timestep_embedding(C.c_void_p(t.data_ptr()), C.c_void_p(out.data_ptr()), t.shape[0], dim,
                                               stream), "gd_timestep_embedding")
    return out


def zero_module(module):
    """nn.py:68-74."""
    for p in module.parameters():
        p.detach().zero_()
    return module
