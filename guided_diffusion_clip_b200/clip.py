"""CLIP ViT image-encoder guidance (BASELINE configs[2]; spec SURVEY §8c).

The reference repository has no CLIP code (only prose, model-card.md:45-48); the architecture is openai/CLIP's visual
transformer as implemented by `transformers.CLIPVisionModelWithProjection` (hidden_act="quick_gelu"), whose
state-dict keys this module uses so that published checkpoints load unchanged.  The guidance is

    cond_fn(x, t) = grad_x  s * < normalize(E(resize_224((x + 1) / 2))), e_txt >

evaluated with hand-written kernels only: fused preprocessing -> patch GEMM -> 12 pre-norm transformer blocks
(LayerNorm, QKV GEMM, masked flash attention over the 197-of-256 padded tokens, projection, QuickGELU MLP) ->
post-LayerNorm of the class token -> projection / cosine head, and the hand-written data-gradient chain back to x
(no parameter gradients, no autograd).  GEMMs run on the tcgen05 implicit-GEMM kernel (gd_conv_igemm, taps = 1)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch as th
import torch.nn as nn

from . import _lib as L
from .engine import Emitter, View, _p, norm_device, pack_1x1, pack_1x1_bwd
from .unet import _register

LN_EPS = 1e-5


class CLIPVisionEncoder(nn.Module):
    """Parameter container (transformers key layout) + plan cache.  Head dimension must be 64 (ViT-B/16: 12 x 64)."""

    def __init__(self, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                 image_size=224, patch_size=16, projection_dim=512):
        super().__init__()
        if hidden_size != 64 * num_attention_heads:
            raise NotImplementedError("CLIP attention head dim != 64 has no CUDA path")
        if hidden_size % 64 or intermediate_size % 64 or (3 * patch_size * patch_size) % 64 or image_size % patch_size:
            raise NotImplementedError("CLIP sizes must be multiples of 64 channels and whole patches")
        self.hidden, self.inter, self.layers, self.heads = hidden_size, intermediate_size, num_hidden_layers, num_attention_heads
        self.image_size, self.patch, self.proj = image_size, patch_size, projection_dim
        self.tokens = 1 + (image_size // patch_size) ** 2
        h, p = hidden_size, "vision_model."

        def reg(name, shape, std):
            _register(self, name, nn.Parameter(th.randn(shape) * std if std else th.zeros(shape)))

        reg(p + "embeddings.class_embedding", (h,), h ** -0.5)
        reg(p + "embeddings.patch_embedding.weight", (h, 3, patch_size, patch_size), 0.02)
        reg(p + "embeddings.position_embedding.weight", (self.tokens, h), 0.02)
        for ln in ("pre_layrnorm",):  # (sic: the transformers key)
            _register(self, p + ln + ".weight", nn.Parameter(th.ones(h)))
            reg(p + ln + ".bias", (h,), 0.0)
        for i in range(num_hidden_layers):
            q = f"{p}encoder.layers.{i}."
            for nm in ("k_proj", "v_proj", "q_proj", "out_proj"):
                reg(q + f"self_attn.{nm}.weight", (h, h), h ** -0.5)
                reg(q + f"self_attn.{nm}.bias", (h,), 0.0)
            _register(self, q + "layer_norm1.weight", nn.Parameter(th.ones(h)))
            reg(q + "layer_norm1.bias", (h,), 0.0)
            reg(q + "mlp.fc1.weight", (intermediate_size, h), h ** -0.5)
            reg(q + "mlp.fc1.bias", (intermediate_size,), 0.0)
            reg(q + "mlp.fc2.weight", (h, intermediate_size), intermediate_size ** -0.5)
            reg(q + "mlp.fc2.bias", (h,), 0.0)
            _register(self, q + "layer_norm2.weight", nn.Parameter(th.ones(h)))
            reg(q + "layer_norm2.bias", (h,), 0.0)
        _register(self, p + "post_layernorm.weight", nn.Parameter(th.ones(h)))
        reg(p + "post_layernorm.bias", (h,), 0.0)
        reg("visual_projection.weight", (projection_dim, h), h ** -0.5)
        self._plans: Dict[tuple, "CLIPPlan"] = {}
        self._param_version = 0
        self.requires_grad_(False)

    def _invalidate(self):
        self._param_version += 1
        self._plans.clear()

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        res = super().load_state_dict(state_dict, strict=strict, **kw)
        self._invalidate()
        return res

    def _apply(self, fn, *a, **kw):
        res = super()._apply(fn, *a, **kw)
        if hasattr(self, "_plans"):
            self._invalidate()
        return res

    def plan(self, n: int, hin: int, win: int, device) -> "CLIPPlan":
        key = (n, hin, win, str(norm_device(device)))
        pl = self._plans.get(key)
        if pl is None:
            pl = CLIPPlan(self, n, hin, win, norm_device(device))
            self._plans[key] = pl
        return pl

    def pooled(self, x: th.Tensor) -> th.Tensor:
        """Post-LayerNorm class-token feature [n, hidden] fp32 of images x in [-1, 1] (before the projection)."""
        n, _, hin, win = x.shape
        return self.plan(n, hin, win, x.device).forward(x).float()


class CLIPPlan:
    """Recorded forward and data-gradient programs for one (batch, input resolution)."""

    LOSS_SCALE = 256.0  # static fp16 gradient scale from the head to the preprocessing transpose

    def __init__(self, model: CLIPVisionEncoder, n: int, hin: int, win: int, device):
        self.model, self.n = model, n
        H, I, heads, T = model.hidden, model.inter, model.heads, model.tokens
        tp = (T + 63) // 64 * 64
        self.t_pad = tp
        dev = device
        fw, bw = Emitter(model, n, dev), Emitter(model, n, dev)
        self.fw, self.bw = fw, bw
        P = fw.P
        pre = "vision_model."
        rows = n * tp
        kdim = 3 * model.patch * model.patch
        self.x_in = th.empty((n, 3, hin, win), dtype=th.float32, device=dev)
        self.text = th.zeros((n, model.proj), dtype=th.float32, device=dev)
        self.sim = th.empty((n,), dtype=th.float32, device=dev)
        self.dx = th.empty((n, 3, hin, win), dtype=th.float32, device=dev)
        self.scale_box = [1.0]

        def tok(c):  # kept activation [n, 1, t_pad, c]
            return fw.act(n, 1, tp, c)

        def ln_fwd(em, x: View, name: str, out: View, stats, nrows=rows, ld=None, ld_out=None):
            em.keep += [stats]
            em.prog.add("gd_layernorm_fwd", C.c_void_p(x.ptr), ld or x.ld, _p(fw.f32(name + ".weight")),
                        _p(fw.f32(name + ".bias")), C.c_float(LN_EPS), C.c_void_p(out.ptr), ld_out or out.ld, _p(stats),
                        nrows, x.c)

        def ln_bwd(x: View, name: str, stats, dy: View, add: Optional[View], dx: View, nrows=rows, ld=None, ld_dy=None,
                   ld_dx=None):
            bw.prog.add("gd_layernorm_bwd", C.c_void_p(x.ptr), ld or x.ld, _p(stats), _p(fw.f32(name + ".weight")),
                        C.c_void_p(dy.ptr), ld_dy or dy.ld, C.c_void_p(add.ptr) if add is not None else None,
                        add.ld if add is not None else 0, C.c_void_p(dx.ptr), ld_dx or dx.ld, nrows, x.c)

        def stats_buf(nrows=rows):
            t = th.empty((nrows, 2), dtype=th.float32, device=dev)
            fw.keep.append(t)
            return t

        # ---------------- forward ----------------
        patches = tok(kdim)
        fw.prog.add("gd_clip_preprocess_fwd", _p(self.x_in), C.c_void_p(patches.ptr), patches.ld, n, hin, win,
                    model.image_size, model.patch, tp)
        # class token + positions as the residual of the patch GEMM (row 0 of `patches` is zero, the conv has no bias)
        pos = th.zeros((n, 1, tp, H), dtype=th.float16, device=dev)
        pe = P[pre + "embeddings.position_embedding.weight"].float()
        pos[:, 0, :T] = pe.to(th.float16)
        pos[:, 0, 0] = (pe[0] + P[pre + "embeddings.class_embedding"].float()).to(th.float16)
        pos_v = View(pos, 0, H)
        fw.keep.append(pos)
        tok0 = tok(H)
        fw.conv(patches, pack_1x1(P[pre + "embeddings.patch_embedding.weight"]), None, H, tok0, taps=1, res=pos_v,
                res_mode=L.RES_SAME)
        st_pre = stats_buf()
        r = tok(H)
        ln_fwd(fw, tok0, pre + "pre_layrnorm", r, st_pre)
        tape = []
        for i in range(model.layers):
            q = f"{pre}encoder.layers.{i}."
            st1, st2 = stats_buf(), stats_buf()
            y = fw.scratch("ln_out", n, 1, tp, H)
            ln_fwd(fw, r, q + "layer_norm1", y, st1)
            wqkv = th.cat([P[q + f"self_attn.{nm}.weight"].float() for nm in ("q_proj", "k_proj", "v_proj")], 0)
            bqkv = th.cat([P[q + f"self_attn.{nm}.bias"].float() for nm in ("q_proj", "k_proj", "v_proj")], 0).contiguous()
            qkv, att = tok(3 * H), tok(H)
            lse = th.empty((n, heads, tp), dtype=th.float32, device=dev)
            fw.keep.append(lse)
            fw.conv(y, pack_1x1(wqkv), bqkv, 3 * H, qkv, taps=1)
            fw.prog.add("gd_attention_fwd_masked", C.c_void_p(qkv.ptr), qkv.ld, C.c_void_p(att.ptr), att.ld, _p(lse), n,
                        tp, T, heads, L.QKV_NEW)
            r2 = tok(H)
            fw.conv(att, pack_1x1(P[q + "self_attn.out_proj.weight"]), fw.f32(q + "self_attn.out_proj.bias"), H, r2,
                    taps=1, res=r, res_mode=L.RES_SAME)
            y2 = fw.scratch("ln_out", n, 1, tp, H)
            ln_fwd(fw, r2, q + "layer_norm2", y2, st2)
            u = tok(I)
            fw.conv(y2, pack_1x1(P[q + "mlp.fc1.weight"]), fw.f32(q + "mlp.fc1.bias"), I, u, taps=1)
            gl = fw.scratch("gelu", n, 1, tp, I)
            fw.prog.add("gd_quickgelu_fwd", C.c_void_p(u.ptr), u.ld, C.c_void_p(gl.ptr), gl.ld, rows, I)
            r3 = tok(H)
            fw.conv(gl, pack_1x1(P[q + "mlp.fc2.weight"]), fw.f32(q + "mlp.fc2.bias"), H, r3, taps=1, res=r2,
                    res_mode=L.RES_SAME)
            tape.append((q, r, st1, wqkv, qkv, att, lse, r2, st2, u))
            r = r3
        # class token rows: row 0 of every image, i.e. n rows with stride t_pad * H
        st_post = stats_buf(n)
        self.fcls = th.empty((n, H), dtype=th.float16, device=dev)
        fcls_v = View(self.fcls.view(n, 1, 1, H), 0, H)
        ln_fwd(fw, r, pre + "post_layernorm", fcls_v, st_post, nrows=n, ld=tp * H, ld_out=H)
        self.r_last = r
        # ---------------- head (similarity + its gradient w.r.t. the pooled feature) ----------------
        self.dfcls = th.empty((n, H), dtype=th.float16, device=dev)
        self.wproj = P["visual_projection.weight"].float().contiguous()
        self._head_args = None  # built per call (the scale is a runtime value)
        # ---------------- backward ----------------
        d_top = th.zeros((n, 1, tp, H), dtype=th.float16, device=dev)  # only the class rows are ever written
        bw.keep.append(d_top)
        d_cur = View(d_top, 0, H)
        dfc_v = View(self.dfcls.view(n, 1, 1, H), 0, H)
        ln_bwd(r, pre + "post_layernorm", st_post, dfc_v, None, d_cur, nrows=n, ld=tp * H, ld_dy=H, ld_dx=tp * H)
        flip = 0
        for (q, r_in, st1, wqkv, qkv, att, lse, r2, st2, u) in reversed(tape):
            d_g = bw.scratch("d_inter", n, 1, tp, I)
            bw.conv(d_cur, pack_1x1_bwd(P[q + "mlp.fc2.weight"]), None, I, d_g, taps=1)
            d_u = bw.scratch("d_inter2", n, 1, tp, I)
            bw.prog.add("gd_quickgelu_bwd", C.c_void_p(u.ptr), u.ld, C.c_void_p(d_g.ptr), d_g.ld, C.c_void_p(d_u.ptr),
                        d_u.ld, rows, I)
            d_y2 = bw.scratch("d_h", n, 1, tp, H)
            bw.conv(d_u, pack_1x1_bwd(P[q + "mlp.fc1.weight"]), None, H, d_y2, taps=1)
            d_r2 = bw.scratch("d_mid", n, 1, tp, H)
            ln_bwd(r2, q + "layer_norm2", st2, d_y2, d_cur, d_r2)
            d_att = bw.scratch("d_h", n, 1, tp, H)
            bw.conv(d_r2, pack_1x1_bwd(P[q + "self_attn.out_proj.weight"]), None, H, d_att, taps=1)
            d_qkv = bw.scratch("d_qkv", n, 1, tp, 3 * H)
            delta = th.empty((n, heads, tp), dtype=th.float32, device=dev)
            bw.keep.append(delta)
            bw.prog.add("gd_attention_bwd_masked", C.c_void_p(qkv.ptr), qkv.ld, C.c_void_p(att.ptr), att.ld,
                        C.c_void_p(d_att.ptr), d_att.ld, _p(lse), _p(delta), C.c_void_p(d_qkv.ptr), d_qkv.ld, n, tp, T,
                        heads, L.QKV_NEW)
            d_y1 = bw.scratch("d_h2", n, 1, tp, H)
            bw.conv(d_qkv, pack_1x1_bwd(wqkv), None, H, d_y1, taps=1)
            flip ^= 1
            d_next = bw.scratch("d_res_a" if flip else "d_res_b", n, 1, tp, H)
            ln_bwd(r_in, q + "layer_norm1", st1, d_y1, d_r2, d_next)
            d_cur = d_next
        d_tok0 = bw.scratch("d_h", n, 1, tp, H)
        ln_bwd(tok0, pre + "pre_layrnorm", st_pre, d_cur, None, d_tok0)
        d_patches = bw.scratch("d_patches", n, 1, tp, kdim)
        bw.conv(d_tok0, pack_1x1_bwd(P[pre + "embeddings.patch_embedding.weight"]), None, kdim, d_patches, taps=1)
        bw.prog.add("gd_clip_preprocess_bwd", C.c_void_p(d_patches.ptr), d_patches.ld, _p(self.dx), n, hin, win,
                    model.image_size, model.patch, tp, C.c_float(1.0 / self.LOSS_SCALE))

    # ---- execution --------------------------------------------------------------------------------------------
    def forward(self, x: th.Tensor) -> th.Tensor:
        """Runs the encoder; returns the post-LayerNorm class-token features [n, hidden] (fp16, plan-owned)."""
        self.x_in.copy_(x)
        self.fw.prog.run()
        return self.fcls

    def _head(self, text: th.Tensor, scale: float, want_grad: bool) -> None:
        m = self.model
        self.text.copy_(text.to(self.text.dtype).expand(self.n, m.proj))
        with th.cuda.device(self.fcls.device):  # launch on the plan's device whatever device is current
            stream = C.c_void_p(th.cuda.current_stream().cuda_stream)
            L.check(L.load().gd_clip_head(_p(self.fcls), m.hidden, _p(self.wproj), _p(self.text), m.proj,
                                          C.c_float(float(scale)), C.c_float(self.LOSS_SCALE), _p(self.sim),
                                          _p(self.dfcls) if want_grad else None, m.hidden, self.n, m.hidden, m.proj,
                                          stream), "gd_clip_head")

    def similarity(self, x: th.Tensor, text: th.Tensor, scale: float) -> th.Tensor:
        self.forward(x)
        self._head(text, scale, False)
        return self.sim

    def guidance(self, x: th.Tensor, text: th.Tensor, scale: float) -> th.Tensor:
        """grad_x sum_b scale * <normalize(E(x_b)), text_b>  (plan-owned fp32 [n,3,h,w])."""
        self.forward(x)
        self._head(text, scale, True)
        self.bw.prog.run()
        return self.dx


class CLIPGuidance:
    """cond_fn(x, t, **kwargs) -> scale * grad_x cos(E_img(x), e_txt): the CLIP text-guidance gradient of
    BASELINE configs[2].  `text_embedding` is a unit vector [proj] (shared) or [n, proj]; `t` is unused (the CLIP
    encoder is not noise-conditioned), exactly like openai/CLIP guided sampling scripts."""

    def __init__(self, encoder: CLIPVisionEncoder, text_embedding: th.Tensor, scale: float = 1.0):
        self.encoder = encoder
        self.text = text_embedding
        self.scale = float(scale)

    def __call__(self, x, t=None, **kwargs):
        n, _, h, w = x.shape
        plan = self.encoder.plan(n, h, w, x.device)
        return plan.guidance(x, self.text.to(x.device), self.scale).clone()
