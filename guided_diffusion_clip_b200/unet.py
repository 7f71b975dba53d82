"""UNetModel / SuperResModel / EncoderUNetModel with the reference's constructor arguments, call signature
and state_dict layout (guided_diffusion/unet.py:396-895) — but no torch compute: a model
here is a *structure spec* plus a parameter tree; `forward` runs a pre-planned sequence of sm_100a
kernels (engine.py) through the C ABI.  torch.nn.Module is used only as the parameter container so that
`load_state_dict / state_dict / to / parameters / eval` behave exactly like the reference objects.

Differences that are deliberate and documented in DESIGN.md:
  * only dims=2 has a CUDA path.  The UNet forward covers both resampling styles (resblock_updown=True: pooled /
    upsampled ResBlocks; False: Downsample / Upsample with conv_resample, unet.py:88-136), both ResBlock
    conditioning styles (FiLM or additive embedding, unet.py:248-255) and head widths 16..256; the classifier
    data-gradient is built for the factory's classifier family (resblock_updown, FiLM, 64-wide heads).
  * activations are always stored fp16 (fp32 accumulate); `use_fp16=False` models run the same kernels.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch as th
import torch.nn as nn


# ------------------------------------------------------------------------------------------------
# structure spec
# ------------------------------------------------------------------------------------------------
@dataclass
class ConvInSpec:
    key: str
    cin: int
    cout: int


@dataclass
class ResSpec:
    key: str
    cin: int
    cout: int
    mode: str  # "none" | "down" | "up"
    film_offset: int = -1  # offset of this block's (scale, shift) in the batched emb_layers output

    @property
    def has_skip_conv(self) -> bool:
        return self.cin != self.cout


@dataclass
class ResampleSpec:
    """Downsample.op (3x3 stride-2 conv, unet.py:125-136) or Upsample.conv after nearest x2 (unet.py:91-110)."""
    key: str   # parameter prefix including ".op" / ".conv"
    ch: int
    mode: str  # "down" | "up"


@dataclass
class AttnSpec:
    key: str
    ch: int
    heads: int
    new_order: bool


@dataclass
class TorsoSpec:
    model_channels: int
    emb_dim: int
    in_channels: int
    input_blocks: List[List[object]] = field(default_factory=list)
    middle_block: List[object] = field(default_factory=list)
    output_blocks: List[List[object]] = field(default_factory=list)
    film_total: int = 0
    out_ch_in: int = 0  # channels entering the `out` head

    def res_blocks(self) -> List[ResSpec]:
        out: List[ResSpec] = []
        for blk in self.input_blocks + [self.middle_block] + self.output_blocks:
            out += [l for l in blk if isinstance(l, ResSpec)]
        return out


class _ParamSpec:
    """Ordered flat list of (dotted name, shape, init kind, is_torso_conv)."""

    def __init__(self):
        self.items: List[Tuple[str, Tuple[int, ...], str, bool]] = []

    def add(self, name: str, shape: Sequence[int], kind: str, torso_conv: bool = False):
        self.items.append((name, tuple(int(s) for s in shape), kind, torso_conv))


def _heads_for(ch: int, num_heads: int, num_head_channels: int) -> int:
    if num_head_channels == -1:
        return num_heads
    if ch % num_head_channels != 0:
        raise AssertionError(
            f"q,k,v channels {ch} is not divisible by num_head_channels {num_head_channels}")
    return ch // num_head_channels


def _add_res(ps: _ParamSpec, key: str, cin: int, cout: int, emb_dim: int, scale_shift: bool, torso: bool):
    ps.add(f"{key}.in_layers.0.weight", (cin,), "ones")
    ps.add(f"{key}.in_layers.0.bias", (cin,), "zeros")
    ps.add(f"{key}.in_layers.2.weight", (cout, cin, 3, 3), "fan_in", torso)
    ps.add(f"{key}.in_layers.2.bias", (cout,), "fan_in:%d" % (cin * 9), torso)
    eo = 2 * cout if scale_shift else cout
    ps.add(f"{key}.emb_layers.1.weight", (eo, emb_dim), "fan_in")
    ps.add(f"{key}.emb_layers.1.bias", (eo,), "fan_in:%d" % emb_dim)
    ps.add(f"{key}.out_layers.0.weight", (cout,), "ones")
    ps.add(f"{key}.out_layers.0.bias", (cout,), "zeros")
    ps.add(f"{key}.out_layers.3.weight", (cout, cout, 3, 3), "zeros", torso)  # zero_module, unet.py:210
    ps.add(f"{key}.out_layers.3.bias", (cout,), "zeros", torso)
    if cin != cout:
        ps.add(f"{key}.skip_connection.weight", (cout, cin, 1, 1), "fan_in", torso)
        ps.add(f"{key}.skip_connection.bias", (cout,), "fan_in:%d" % cin, torso)


def _add_resample(ps: _ParamSpec, key: str, ch: int):
    ps.add(f"{key}.weight", (ch, ch, 3, 3), "fan_in", True)
    ps.add(f"{key}.bias", (ch,), "fan_in:%d" % (ch * 9), True)


def _add_attn(ps: _ParamSpec, key: str, ch: int, torso: bool):
    ps.add(f"{key}.norm.weight", (ch,), "ones")
    ps.add(f"{key}.norm.bias", (ch,), "zeros")
    ps.add(f"{key}.qkv.weight", (3 * ch, ch, 1), "fan_in", torso)
    ps.add(f"{key}.qkv.bias", (3 * ch,), "fan_in:%d" % ch, torso)
    ps.add(f"{key}.proj_out.weight", (ch, ch, 1), "zeros", torso)  # zero_module, unet.py:294
    ps.add(f"{key}.proj_out.bias", (ch,), "zeros", torso)


def build_torso(
    *, in_channels: int, model_channels: int, num_res_blocks: int, attention_resolutions, channel_mult,
    num_heads: int, num_head_channels: int, num_heads_upsample: int, use_scale_shift_norm: bool,
    resblock_updown: bool, use_new_attention_order: bool, decoder: bool, ps: _ParamSpec,
    conv_resample: bool = True,
) -> TorsoSpec:
    """Walk the same construction order as unet.py:481-611 (UNet) / :739-823 (encoder)."""
    if not resblock_updown and not conv_resample and len(channel_mult) > 1:
        raise NotImplementedError(
            "conv_resample=False (parameter-free AvgPool2d / nearest resampling layers between levels) has no CUDA "
            "path; no factory of the reference passes it (script_util.py:130-167)")
    emb_dim = model_channels * 4
    spec = TorsoSpec(model_channels=model_channels, emb_dim=emb_dim, in_channels=in_channels)
    ch = int(channel_mult[0] * model_channels)
    ps.add("input_blocks.0.0.weight", (ch, in_channels, 3, 3), "fan_in", True)
    ps.add("input_blocks.0.0.bias", (ch,), "fan_in:%d" % (in_channels * 9), True)
    spec.input_blocks.append([ConvInSpec("input_blocks.0.0", in_channels, ch)])
    chans = [ch]
    ds = 1
    for level, mult in enumerate(channel_mult):
        for _ in range(num_res_blocks):
            idx = len(spec.input_blocks)
            cout = int(mult * model_channels)
            layers: List[object] = [ResSpec(f"input_blocks.{idx}.0", ch, cout, "none")]
            _add_res(ps, layers[0].key, ch, cout, emb_dim, use_scale_shift_norm, True)
            ch = cout
            if ds in attention_resolutions:
                a = AttnSpec(f"input_blocks.{idx}.1", ch, _heads_for(ch, num_heads, num_head_channels),
                             use_new_attention_order)
                _add_attn(ps, a.key, ch, True)
                layers.append(a)
            spec.input_blocks.append(layers)
            chans.append(ch)
        if level != len(channel_mult) - 1:
            idx = len(spec.input_blocks)
            if resblock_updown:
                r = ResSpec(f"input_blocks.{idx}.0", ch, ch, "down")
                _add_res(ps, r.key, ch, ch, emb_dim, use_scale_shift_norm, True)
            else:
                r = ResampleSpec(f"input_blocks.{idx}.0.op", ch, "down")
                _add_resample(ps, r.key, ch)
            spec.input_blocks.append([r])
            chans.append(ch)
            ds *= 2
    mid = [ResSpec("middle_block.0", ch, ch, "none"),
           AttnSpec("middle_block.1", ch, _heads_for(ch, num_heads, num_head_channels), use_new_attention_order),
           ResSpec("middle_block.2", ch, ch, "none")]
    _add_res(ps, mid[0].key, ch, ch, emb_dim, use_scale_shift_norm, True)
    _add_attn(ps, mid[1].key, ch, True)
    _add_res(ps, mid[2].key, ch, ch, emb_dim, use_scale_shift_norm, True)
    spec.middle_block = mid
    if decoder:
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                ich = chans.pop()
                idx = len(spec.output_blocks)
                cout = int(model_channels * mult)
                layers = [ResSpec(f"output_blocks.{idx}.0", ch + ich, cout, "none")]
                _add_res(ps, layers[0].key, ch + ich, cout, emb_dim, use_scale_shift_norm, True)
                ch = cout
                if ds in attention_resolutions:
                    a = AttnSpec(f"output_blocks.{idx}.{len(layers)}", ch,
                                 _heads_for(ch, num_heads_upsample, num_head_channels), use_new_attention_order)
                    _add_attn(ps, a.key, ch, True)
                    layers.append(a)
                if level and i == num_res_blocks:
                    if resblock_updown:
                        r = ResSpec(f"output_blocks.{idx}.{len(layers)}", ch, ch, "up")
                        _add_res(ps, r.key, ch, ch, emb_dim, use_scale_shift_norm, True)
                    else:
                        r = ResampleSpec(f"output_blocks.{idx}.{len(layers)}.conv", ch, "up")
                        _add_resample(ps, r.key, ch)
                    layers.append(r)
                    ds //= 2
                spec.output_blocks.append(layers)
    spec.out_ch_in = ch
    off = 0
    for r in spec.res_blocks():
        r.film_offset = off
        off += 2 * r.cout if use_scale_shift_norm else r.cout
    spec.film_total = off
    return spec


# ------------------------------------------------------------------------------------------------
# parameter container
# ------------------------------------------------------------------------------------------------
def _init_param(shape, kind: str) -> th.Tensor:
    if kind == "zeros":
        return th.zeros(shape)
    if kind == "ones":
        return th.ones(shape)
    if kind == "normal":
        return th.randn(shape)
    if kind.startswith("pos:"):
        return th.randn(shape) / float(kind.split(":")[1]) ** 0.5
    if kind.startswith("fan_in"):
        fan_in = int(kind.split(":")[1]) if ":" in kind else int(math.prod(shape[1:]))
        bound = 1.0 / math.sqrt(fan_in)
        return th.empty(shape).uniform_(-bound, bound)
    raise ValueError(kind)


def _register(root: nn.Module, dotted: str, p: nn.Parameter) -> None:
    parts = dotted.split(".")
    m = root
    for part in parts[:-1]:
        if part not in m._modules:
            m.add_module(part, nn.Module())
        m = m._modules[part]
    m.register_parameter(parts[-1], p)


class _GdModule(nn.Module):
    """Parameter tree + plan cache shared by the three model classes."""

    def _materialise(self, ps: _ParamSpec) -> None:
        self._torso_conv_names = [name for name, _, _, tc in ps.items if tc]
        for name, shape, kind, _ in ps.items:
            _register(self, name, nn.Parameter(_init_param(shape, kind)))
        self._plans: Dict[tuple, object] = {}
        self._param_version = 0

    # ---- reference protocol -----------------------------------------------------------------
    def _cast_torso(self, dtype) -> None:
        params = dict(self.named_parameters())
        for name in self._torso_conv_names:
            params[name].data = params[name].data.to(dtype)
        self._invalidate()

    def convert_to_fp16(self):
        """fp16_util.py:15-22 applied to the torso (unet.py:619-625, 858-863): conv weights+biases only."""
        self._cast_torso(th.float16)

    def convert_to_fp32(self):
        self._cast_torso(th.float32)

    def _invalidate(self) -> None:
        self._param_version += 1
        self._plans.clear()

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        res = super().load_state_dict(state_dict, strict=strict, **kw)
        self._invalidate()
        return res

    def _apply(self, fn, *a, **kw):  # .to() / .cuda() / .half()
        res = super()._apply(fn, *a, **kw)
        if hasattr(self, "_plans"):
            self._invalidate()
        return res

    def _plan_for(self, key: tuple, builder):
        plan = self._plans.get(key)
        if plan is None:
            plan = builder()
            self._plans[key] = plan
        return plan


class UNetModel(_GdModule):
    """Drop-in for guided_diffusion.unet.UNetModel (unet.py:396-664).

    `label_mlp=True` selects the fork's CLIP-feature conditioning (unet_other.py:25-41): label_emb is
    Linear(num_classes,4C)-SiLU-Linear(4C,4C) fed by `clip_feat` instead of nn.Embedding fed by `y`.
    """

    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks,
                 attention_resolutions, dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2,
                 num_classes=None, use_checkpoint=False, use_fp16=False, num_heads=1, num_head_channels=-1,
                 num_heads_upsample=-1, use_scale_shift_norm=False, resblock_updown=False,
                 use_new_attention_order=False, label_mlp=False):
        super().__init__()
        if dims != 2:
            raise NotImplementedError("only dims=2 is on the sampling path (unet.py:102-105,129 are out of scope)")
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads
        self.image_size = image_size
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = tuple(attention_resolutions)
        self.dropout = dropout
        self.channel_mult = tuple(channel_mult)
        self.conv_resample = conv_resample
        self.num_classes = num_classes
        self.use_checkpoint = use_checkpoint
        self.dtype = th.float16 if use_fp16 else th.float32
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels
        self.num_heads_upsample = num_heads_upsample
        self.use_scale_shift_norm = use_scale_shift_norm
        self.label_mlp = bool(label_mlp)
        self.time_embed_dim = model_channels * 4

        ps = _ParamSpec()
        e = self.time_embed_dim
        ps.add("time_embed.0.weight", (e, model_channels), "fan_in")
        ps.add("time_embed.0.bias", (e,), "fan_in:%d" % model_channels)
        ps.add("time_embed.2.weight", (e, e), "fan_in")
        ps.add("time_embed.2.bias", (e,), "fan_in:%d" % e)
        if num_classes is not None:
            if self.label_mlp:
                ps.add("label_emb.0.weight", (e, num_classes), "fan_in")
                ps.add("label_emb.0.bias", (e,), "fan_in:%d" % num_classes)
                ps.add("label_emb.2.weight", (e, e), "fan_in")
                ps.add("label_emb.2.bias", (e,), "fan_in:%d" % e)
            else:
                ps.add("label_emb.weight", (num_classes, e), "normal")
        self.spec = build_torso(
            in_channels=in_channels, model_channels=model_channels, num_res_blocks=num_res_blocks,
            attention_resolutions=self.attention_resolutions, channel_mult=self.channel_mult, num_heads=num_heads,
            num_head_channels=num_head_channels, num_heads_upsample=num_heads_upsample,
            use_scale_shift_norm=use_scale_shift_norm, resblock_updown=resblock_updown,
            use_new_attention_order=use_new_attention_order, decoder=True, ps=ps, conv_resample=conv_resample)
        ps.add("out.0.weight", (self.spec.out_ch_in,), "ones")
        ps.add("out.0.bias", (self.spec.out_ch_in,), "zeros")
        ps.add("out.2.weight", (out_channels, self.spec.out_ch_in, 3, 3), "zeros")  # zero_module, unet.py:616
        ps.add("out.2.bias", (out_channels,), "zeros")
        self._extra_params(ps)
        self._materialise(ps)

    def _extra_params(self, ps: _ParamSpec) -> None:
        pass

    # -- checkpoints of the other label variant ---------------------------------------------------
    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        """The reference fork's factory always builds the CLIP-feature variant (label_emb = Linear-SiLU-Linear over a
        512-d feature, script_util.py:168, unet_other.py:25-41) while upstream checkpoints carry an nn.Embedding table.
        A checkpoint of the OTHER variant than the one this model was built with is adopted: the label parameters are
        re-created in the checkpoint's layout (and the model then expects `clip_feat` resp. integer `y`), so
        reference-trained weights load with strict=True whichever `conditioning` the factory was called with."""
        if self.num_classes is not None:
            ck_mlp, ck_emb = "label_emb.0.weight" in state_dict, "label_emb.weight" in state_dict
            if (ck_mlp and not self.label_mlp) or (ck_emb and self.label_mlp):
                if ("bias_feat" in state_dict) != hasattr(self, "bias_feat"):
                    raise RuntimeError(
                        "checkpoint and model are different super-resolution variants (SRImageModel_Feat carries "
                        "`bias_feat` and takes clip_feat / clip_feat2 / img2; SuperResModel takes low_res): build the "
                        "model with sr_create_model(..., conditioning='clip_feat' | 'low_res') to match the checkpoint")
                self._adopt_label_variant(state_dict, ck_mlp)
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def _adopt_label_variant(self, state_dict, to_mlp: bool) -> None:
        ref = next(self.parameters())
        self._modules.pop("label_emb", None)
        names = (["label_emb.0.weight", "label_emb.0.bias", "label_emb.2.weight", "label_emb.2.bias"] if to_mlp
                 else ["label_emb.weight"])
        for nm in names:
            _register(self, nm, nn.Parameter(th.zeros(tuple(state_dict[nm].shape), dtype=th.float32, device=ref.device)))
        self.label_mlp = to_mlp
        self.num_classes = int(state_dict[names[0]].shape[1 if to_mlp else 0])
        self._invalidate()

    # -- conditioning vector -------------------------------------------------------------------
    def _cond(self, y, kwargs):
        feat = kwargs.get("clip_feat")
        if self.label_mlp and feat is not None:  # UNetModel_clip_feat.forward (unet_other.py:36-41)
            return feat.squeeze().float().reshape(-1, self.num_classes)
        return y

    def _check_labels(self, y) -> None:
        """nn.Embedding raises IndexError for labels outside [0, num_classes); the gather kernel cannot, so labels are
        validated on the host once per distinct tensor (not while a CUDA graph is being captured)."""
        key = (y.data_ptr(), y._version, tuple(y.shape))
        if getattr(self, "_labels_ok", None) == key or (y.is_cuda and th.cuda.is_current_stream_capturing()):
            return
        if y.dtype not in (th.int64, th.int32, th.int16, th.uint8, th.int8):
            raise TypeError(
                f"this model embeds integer class labels (nn.Embedding, unet.py:478-479) but got y of dtype {y.dtype}; "
                "for the fork's CLIP-feature conditioning build it with conditioning='clip_feat' or load a fork checkpoint")
        lo, hi = int(y.min()), int(y.max())
        if lo < 0 or hi >= self.num_classes:
            raise IndexError(f"class label out of range: y in [{lo}, {hi}], num_classes = {self.num_classes}")
        self._labels_ok = key

    def forward(self, x, timesteps, y=None, **kwargs):
        """x: [N, C, H, W] fp32; timesteps: [N]; y: [N] int64 labels (or [N,num_classes] features when
        label_mlp).  Returns [N, out_channels, H, W] fp32 (unet.py:635-664)."""
        y = self._cond(y, kwargs)
        if y is None and self.num_classes is not None and not self.label_mlp and kwargs.get("clip_feat") is not None:
            raise TypeError("clip_feat was passed to a model built for integer class labels (conditioning='labels'); "
                            "build it with conditioning='clip_feat' or load a fork checkpoint (label_emb.0/2.*)")
        assert (y is not None) == (self.num_classes is not None), \
            "must specify y if and only if the model is class-conditional"
        if y is not None:
            assert y.shape[0] == x.shape[0], f"{y.shape} != {x.shape}"
            if not self.label_mlp:
                self._check_labels(y)
            if y.device != x.device:
                raise RuntimeError(f"y lives on {y.device} but x on {x.device}")
        from .engine import UNetPlan, norm_device
        n, c, h, w = x.shape
        assert c == self.in_channels, f"expected {self.in_channels} input channels, got {c}"
        dev = norm_device(x.device)
        plan = self._plan_for(("unet", n, h, w, str(dev)), lambda: UNetPlan(self, n, h, w, dev))
        return plan.run(x, timesteps, y).clone()


class SuperResModel(UNetModel):
    """unet.py:667-681: bilinear-upsampled `low_res` is concatenated to x (in_channels doubles)."""

    def __init__(self, image_size, in_channels, *args, **kwargs):
        super().__init__(image_size, in_channels * 2, *args, **kwargs)
        self._base_in = in_channels

    def forward(self, x, timesteps, low_res=None, **kwargs):
        from .engine import bilinear_concat
        assert low_res is not None, "SuperResModel needs the low_res kwarg"
        return super().forward(bilinear_concat(x, low_res), timesteps, **kwargs)


class UNetModel_clip_feat(UNetModel):
    """Fork variant returned by the reference factory (unet_other.py:25-41, script_util.py:168)."""

    def __init__(self, image_size, in_channels, *args, **kwargs):
        kwargs["label_mlp"] = True
        super().__init__(image_size, in_channels, *args, **kwargs)



class SRImageModel_Feat(UNetModel):
    """Fork super-res variant (unet_other.py:43-77): x ‖ img2 as input, y = clip_feat - clip_feat2 + bias_feat."""

    def __init__(self, image_size, in_channels, *args, **kwargs):
        kwargs["label_mlp"] = True
        super().__init__(image_size, in_channels * 2, *args, **kwargs)

    def _extra_params(self, ps: _ParamSpec) -> None:
        if self.num_classes is not None:
            ps.add("bias_feat", (self.num_classes,), "normal")

    def forward(self, x, timesteps, clip_feat=None, clip_feat2=None, img2=None, **kwargs):
        y = clip_feat.squeeze().float().reshape(-1, self.num_classes)
        y2 = clip_feat2.squeeze().float().reshape(-1, self.num_classes)
        y = y - y2 + self.bias_feat.to(y.device)
        return super().forward(th.cat([x, img2], dim=1), timesteps, y=y, **kwargs)


class EncoderUNetModel(_GdModule):
    """Drop-in for guided_diffusion.unet.EncoderUNetModel with pool="attention" (unet.py:684-895).

    Calling it returns logits [N, out_channels].  The call is autograd-aware: if `x.requires_grad`, the
    returned logits carry a grad_fn whose backward runs the hand-written dX kernels, so the reference's own
    cond_fn closure (scripts/classifier_sample.py:54-61) works unmodified.
    """

    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks,
                 attention_resolutions, dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2,
                 use_checkpoint=False, use_fp16=False, num_heads=1, num_head_channels=-1, num_heads_upsample=-1,
                 use_scale_shift_norm=False, resblock_updown=False, use_new_attention_order=False,
                 pool="adaptive"):
        super().__init__()
        if dims != 2:
            raise NotImplementedError("only dims=2 is on the sampling path")
        if pool != "attention":
            raise NotImplementedError(
                f"pool={pool!r}: only the attention pool is built (the factory never passes anything else, "
                "script_util.py:40,268)")
        assert num_head_channels != -1
        self.image_size = image_size
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = tuple(attention_resolutions)
        self.channel_mult = tuple(channel_mult)
        self.dtype = th.float16 if use_fp16 else th.float32
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels
        self.use_scale_shift_norm = use_scale_shift_norm
        self.pool = pool
        ps = _ParamSpec()
        e = model_channels * 4
        ps.add("time_embed.0.weight", (e, model_channels), "fan_in")
        ps.add("time_embed.0.bias", (e,), "fan_in:%d" % model_channels)
        ps.add("time_embed.2.weight", (e, e), "fan_in")
        ps.add("time_embed.2.bias", (e,), "fan_in:%d" % e)
        self.spec = build_torso(
            in_channels=in_channels, model_channels=model_channels, num_res_blocks=num_res_blocks,
            attention_resolutions=self.attention_resolutions, channel_mult=self.channel_mult, num_heads=num_heads,
            num_head_channels=num_head_channels, num_heads_upsample=num_heads_upsample,
            use_scale_shift_norm=use_scale_shift_norm, resblock_updown=resblock_updown,
            use_new_attention_order=use_new_attention_order, decoder=False, ps=ps, conv_resample=conv_resample)
        ch = self.spec.out_ch_in
        ds = 2 ** (len(self.channel_mult) - 1)
        self.pool_spatial = image_size // ds
        self.pool_heads = ch // num_head_channels
        ps.add("out.0.weight", (ch,), "ones")
        ps.add("out.0.bias", (ch,), "zeros")
        ps.add("out.2.positional_embedding", (ch, self.pool_spatial ** 2 + 1), "pos:%d" % ch)
        ps.add("out.2.qkv_proj.weight", (3 * ch, ch, 1), "fan_in")
        ps.add("out.2.qkv_proj.bias", (3 * ch,), "fan_in:%d" % ch)
        ps.add("out.2.c_proj.weight", (out_channels, ch, 1), "fan_in")
        ps.add("out.2.c_proj.bias", (out_channels,), "fan_in:%d" % ch)
        self._materialise(ps)

    def plan(self, n: int, h: int, w: int, device):
        from .engine import ClassifierPlan, norm_device
        dev = norm_device(device)
        return self._plan_for(("clf", n, h, w, str(dev)), lambda: ClassifierPlan(self, n, h, w, dev))

    def forward(self, x, timesteps):
        from .engine import classifier_apply
        return classifier_apply(self, x, timesteps)
