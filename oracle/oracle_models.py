"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product package (guided_diffusion_clip_b200);
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.

CPU restatement, in plain functional PyTorch fp32, of the reference's model forward passes:
  UNetModel.forward            /root/reference/guided_diffusion/unet.py:635-664
  SuperResModel.forward        unet.py:677-681
  EncoderUNetModel.forward     unet.py:872-895  (pool="attention": AttentionPool2d.forward unet.py:43-51)
  ResBlock._forward            unet.py:236-256
  AttentionBlock._forward      unet.py:299-305, QKVAttentionLegacy :337-354, QKVAttention :370-389
  GroupNorm32 / timestep_embedding   nn.py:17-19, 103-121
It works directly on a reference-layout state_dict and derives the block structure from the
keys, so it shares no code with the product's planner.  Parity pinning: oracle/make_golden.py runs the REAL
reference modules (imported from /root/reference in the build container) on the same state_dict and inputs
and commits their outputs under tests/golden/; tests/test_oracle_golden.py checks this file against them.
"""
from __future__ import annotations

import math
import re
from typing import Dict, List, Optional

import torch as th
import torch.nn.functional as F


def timestep_embedding(t: th.Tensor, dim: int, max_period: float = 10000.0) -> th.Tensor:
    half = dim // 2
    freqs = th.exp(-math.log(max_period) * th.arange(half, dtype=th.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    emb = th.cat([th.cos(args), th.sin(args)], dim=-1)
    if dim % 2:
        emb = th.cat([emb, th.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _gn(x: th.Tensor, sd, key: str) -> th.Tensor:
    return F.group_norm(x.float(), 32, sd[key + ".weight"].float(), sd[key + ".bias"].float(), eps=1e-5)


def _conv(x, sd, key, padding):
    w = sd[key + ".weight"].float()
    if w.dim() == 3:
        w = w[..., None]
    return F.conv2d(x, w, sd[key + ".bias"].float(), padding=padding)


def res_block(x, emb, sd, key: str, mode: str):
    h = F.silu(_gn(x, sd, key + ".in_layers.0"))
    if mode == "down":
        h, x = F.avg_pool2d(h, 2), F.avg_pool2d(x, 2)
    elif mode == "up":
        h, x = F.interpolate(h, scale_factor=2, mode="nearest"), F.interpolate(x, scale_factor=2, mode="nearest")
    h = _conv(h, sd, key + ".in_layers.2", 1)
    e = F.linear(F.silu(emb), sd[key + ".emb_layers.1.weight"].float(), sd[key + ".emb_layers.1.bias"].float())
    cout = h.shape[1]
    if e.shape[1] == 2 * cout:  # use_scale_shift_norm (unet.py:248-252)
        scale, shift = e[:, :cout, None, None], e[:, cout:, None, None]
        h = _gn(h, sd, key + ".out_layers.0") * (1 + scale) + shift
    else:                       # additive embedding, then the plain out_layers (unet.py:253-255)
        assert e.shape[1] == cout
        h = _gn(h + e[:, :, None, None], sd, key + ".out_layers.0")
    h = _conv(F.silu(h), sd, key + ".out_layers.3", 1)
    if key + ".skip_connection.weight" in sd:
        x = _conv(x, sd, key + ".skip_connection", 0)
    return x + h


def qkv_attention(qkv: th.Tensor, heads: int, new_order: bool) -> th.Tensor:
    """qkv [N, 3*H*d, T] -> [N, H*d, T]."""
    n, width, t = qkv.shape
    d = width // (3 * heads)
    if new_order:
        q, k, v = qkv.chunk(3, dim=1)
        q, k, v = (z.reshape(n * heads, d, t) for z in (q, k, v))
    else:
        q, k, v = qkv.reshape(n * heads, 3 * d, t).split(d, dim=1)
    s = 1.0 / math.sqrt(math.sqrt(d))
    w = th.einsum("bct,bcs->bts", q * s, k * s)
    w = th.softmax(w.float(), dim=-1)
    return th.einsum("bts,bcs->bct", w, v).reshape(n, -1, t)


def attn_block(x, sd, key: str, head_dim: int, new_order: bool):
    """head_dim > 0: num_head_channels; head_dim < 0: a fixed head count -head_dim (num_head_channels=-1, unet.py:279-285)."""
    n, c, hh, ww = x.shape
    g = _gn(x, sd, key + ".norm")
    qkv = _conv(g, sd, key + ".qkv", 0).reshape(n, 3 * c, hh * ww)
    a = qkv_attention(qkv, c // head_dim if head_dim > 0 else -head_dim, new_order).reshape(n, c, hh, ww)
    return x + _conv(a, sd, key + ".proj_out", 0)


def _block_layers(sd, prefix: str) -> List[str]:
    idx = sorted({int(m.group(1)) for k in sd for m in [re.match(re.escape(prefix) + r"\.(\d+)\.", k)] if m})
    return [f"{prefix}.{i}" for i in idx]


def _run_layers(h, emb, sd, layer_keys: List[str], head_dim: int, new_order: bool, modes: Dict[str, str]):
    for lk in layer_keys:
        if lk + ".qkv.weight" in sd:
            h = attn_block(h, sd, lk, head_dim, new_order)
        elif lk + ".in_layers.0.weight" in sd:
            h = res_block(h, emb, sd, lk, modes.get(lk, "none"))
        elif lk + ".op.weight" in sd:    # Downsample with conv_resample: 3x3 stride 2 (unet.py:125-136)
            h = F.conv2d(h, sd[lk + ".op.weight"].float(), sd[lk + ".op.bias"].float(), stride=2, padding=1)
        elif lk + ".conv.weight" in sd:  # Upsample: nearest x2, then 3x3 conv (unet.py:100-110)
            h = _conv(F.interpolate(h, scale_factor=2, mode="nearest"), sd, lk + ".conv", 1)
        else:
            h = _conv(h, sd, lk, 1)
    return h


def _n_blocks(sd, prefix: str) -> int:
    return 1 + max(int(m.group(1)) for k in sd for m in [re.match(re.escape(prefix) + r"\.(\d+)\.", k)] if m)


def infer_modes(sd, channel_mult_len: int, num_res_blocks: int, decoder: bool) -> Dict[str, str]:
    """Which ResBlocks resample (resblock_updown=True layout of unet.py:515-537, 593-609)."""
    modes: Dict[str, str] = {}
    idx = 1
    for level in range(channel_mult_len):
        idx += num_res_blocks
        if level != channel_mult_len - 1:
            modes[f"input_blocks.{idx}.0"] = "down"
            idx += 1
    if decoder:
        j = 0
        for level in reversed(range(channel_mult_len)):
            for i in range(num_res_blocks + 1):
                if level and i == num_res_blocks:
                    last = _block_layers(sd, f"output_blocks.{j}")[-1]
                    modes[last] = "up"
                j += 1
    return modes


def _embed(sd, t, y, model_channels):
    emb = timestep_embedding(t, model_channels)
    emb = F.linear(emb, sd["time_embed.0.weight"].float(), sd["time_embed.0.bias"].float())
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"].float(), sd["time_embed.2.bias"].float())
    if y is not None:
        if "label_emb.weight" in sd:
            emb = emb + sd["label_emb.weight"].float()[y]
        else:  # fork MLP, unet_other.py:29-33
            l = F.linear(y.float(), sd["label_emb.0.weight"].float(), sd["label_emb.0.bias"].float())
            emb = emb + F.linear(F.silu(l), sd["label_emb.2.weight"].float(), sd["label_emb.2.bias"].float())
    return emb


def unet_forward(sd, x, t, y=None, *, num_res_blocks: int, channel_mult_len: int, head_dim: int = 64,
                 new_order: bool = False, low_res: Optional[th.Tensor] = None, num_heads: int = 0,
                 num_heads_upsample: int = 0) -> th.Tensor:
    """num_heads > 0 selects num_head_channels=-1 semantics: a fixed head count (num_heads_upsample in the decoder)."""
    sd = {k: v.float() for k, v in sd.items()}
    hd_up = head_dim
    if num_heads > 0:
        head_dim, hd_up = -num_heads, -(num_heads_upsample or num_heads)
    if low_res is not None:
        x = th.cat([x, F.interpolate(low_res, x.shape[2:], mode="bilinear")], dim=1)
    mc = sd["time_embed.0.weight"].shape[1]
    emb = _embed(sd, t, y, mc)
    modes = infer_modes(sd, channel_mult_len, num_res_blocks, True)
    hs = []
    h = x.float()
    for i in range(_n_blocks(sd, "input_blocks")):
        h = _run_layers(h, emb, sd, _block_layers(sd, f"input_blocks.{i}"), head_dim, new_order, modes)
        hs.append(h)
    h = _run_layers(h, emb, sd, _block_layers(sd, "middle_block"), head_dim, new_order, modes)
    for j in range(_n_blocks(sd, "output_blocks")):
        h = th.cat([h, hs.pop()], dim=1)
        h = _run_layers(h, emb, sd, _block_layers(sd, f"output_blocks.{j}"), hd_up, new_order, modes)
    h = F.silu(_gn(h, sd, "out.0"))
    return _conv(h, sd, "out.2", 1)


def classifier_forward(sd, x, t, *, num_res_blocks: int, channel_mult_len: int, head_dim: int = 64) -> th.Tensor:
    sd = {k: v.float() for k, v in sd.items()}
    mc = sd["time_embed.0.weight"].shape[1]
    emb = _embed(sd, t, None, mc)
    modes = infer_modes(sd, channel_mult_len, num_res_blocks, False)
    h = x.float()
    for i in range(_n_blocks(sd, "input_blocks")):
        h = _run_layers(h, emb, sd, _block_layers(sd, f"input_blocks.{i}"), head_dim, False, modes)
    h = _run_layers(h, emb, sd, _block_layers(sd, "middle_block"), head_dim, False, modes)
    h = F.silu(_gn(h, sd, "out.0"))
    n, c = h.shape[:2]
    tok = h.reshape(n, c, -1)
    tok = th.cat([tok.mean(dim=-1, keepdim=True), tok], dim=-1) + sd["out.2.positional_embedding"][None]
    qkv = F.conv1d(tok, sd["out.2.qkv_proj.weight"], sd["out.2.qkv_proj.bias"])
    a = qkv_attention(qkv, c // head_dim, True)
    return F.conv1d(a, sd["out.2.c_proj.weight"], sd["out.2.c_proj.bias"])[:, :, 0]


def classifier_guidance(sd, x, t, y, scale: float, **kw) -> th.Tensor:
    """cond_fn of scripts/classifier_sample.py:54-61."""
    with th.enable_grad():
        x_in = x.detach().float().requires_grad_(True)
        logits = classifier_forward(sd, x_in, t, **kw)
        logp = F.log_softmax(logits, dim=-1)
        sel = logp[range(len(logits)), y.view(-1)]
        return th.autograd.grad(sel.sum(), x_in)[0] * scale


def make_state_dict(shapes: Dict[str, tuple], seed: int) -> Dict[str, th.Tensor]:
    """Deterministic random weights in the reference key layout (nothing left at its zero_module init)."""
    g = th.Generator().manual_seed(seed)
    sd = {}
    for name, shape in shapes.items():
        shape = tuple(shape)
        is_norm = bool(re.search(r"(in_layers\.0|out_layers\.0|\.norm|^out\.0)\.(weight|bias)$", name))
        if is_norm and name.endswith("weight"):
            v = 1.0 + 0.1 * th.randn(shape, generator=g)
        elif len(shape) == 1:
            v = 0.1 * th.randn(shape, generator=g)
        elif name.endswith("positional_embedding"):
            v = th.randn(shape, generator=g) / shape[0] ** 0.5
        elif name == "label_emb.weight":
            v = 0.5 * th.randn(shape, generator=g)
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            v = th.randn(shape, generator=g) / math.sqrt(fan_in)
        sd[name] = v
    return sd


def srfeat_forward(sd, x, t, clip_feat, clip_feat2, img2, **kw) -> th.Tensor:
    """SRImageModel_Feat.forward (unet_other.py:56-77): y = clip_feat - clip_feat2 + bias_feat feeds the label MLP,
    the reference image img2 is concatenated to x_t."""
    n = x.shape[0]
    y = clip_feat.reshape(n, -1).float() - clip_feat2.reshape(n, -1).float() + sd["bias_feat"].float()
    return unet_forward(sd, th.cat([x, img2], dim=1), t, y, **kw)
