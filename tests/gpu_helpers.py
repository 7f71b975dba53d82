"""Thin test-side wrappers that call the C ABI (include/gd_b200.h) on torch CUDA tensors."""
from __future__ import annotations

import ctypes as C

import torch as th

from guided_diffusion_clip_b200 import _lib as L


def stream():
    return C.c_void_p(th.cuda.current_stream().cuda_stream)


def vp(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def nhwc_half(x_nchw: th.Tensor, ld: int = None, off: int = 0) -> th.Tensor:
    """NCHW float -> NHWC half buffer with per-pixel stride ld (channels placed at [off, off+C))."""
    n, c, h, w = x_nchw.shape
    ld = ld or c
    buf = th.zeros((n, h, w, ld), dtype=th.float16, device=x_nchw.device)
    buf[..., off:off + c] = x_nchw.permute(0, 2, 3, 1).half()
    return buf


def conv_igemm(a0_buf, c0, off0, wpack, bias, cout, n, h, w, *, taps=9, a1_buf=None, c1=0, off1=0, res_buf=None,
               res_off=0, res_mode=L.RES_NONE, out_mode=L.OUT_NHWC_F16, ld_out=None, out_off=0, out_scale=1.0, bn=0,
               stats_out=None, out_buf=None, gn=None, splitk_ws=None):
    """gn = dict(mode, silu, coef): GroupNorm fused into the main operand (a0 is the raw tensor; coef from gn_coef)."""
    lib = L.load()
    d = L.ConvDesc()
    d.a0, d.c0, d.ld0, d.taps = a0_buf.data_ptr() + 2 * off0, c0, a0_buf.shape[-1], taps
    if a1_buf is not None:
        d.a1, d.c1, d.ld1 = a1_buf.data_ptr() + 2 * off1, c1, a1_buf.shape[-1]
    d.n, d.h, d.w = n, h, w
    d.wpack, d.k_total, d.n_pad = wpack.data_ptr(), wpack.shape[1], wpack.shape[0]
    d.bias = bias.data_ptr() if bias is not None else None
    d.cout = cout
    if res_buf is not None:
        d.res, d.ld_res, d.res_mode = res_buf.data_ptr() + 2 * res_off, res_buf.shape[-1], res_mode
    if out_mode == L.OUT_NHWC_F16:
        ld_out = ld_out or cout
        out = out_buf if out_buf is not None else th.zeros((n, h, w, ld_out), dtype=th.float16, device=a0_buf.device)
        d.out, d.ld_out = out.data_ptr() + 2 * out_off, ld_out
    else:
        out = th.zeros((n, cout, h, w), dtype=th.float32, device=a0_buf.device)
        d.out, d.ld_out = out.data_ptr(), 0
    d.out_mode, d.bn, d.out_scale = out_mode, bn, out_scale
    d.stats_out = stats_out.data_ptr() if stats_out is not None else None
    if gn is not None:
        d.gn_mode, d.gn_silu = gn["mode"], int(gn.get("silu", True))
        d.gn_coef = gn["coef"].data_ptr()
    if splitk_ws is not None:
        d.splitk_ws, d.splitk_ws_bytes = splitk_ws.data_ptr(), splitk_ws.numel() * splitk_ws.element_size()
    L.check(lib.gd_conv_igemm(C.byref(d), stream()), "gd_conv_igemm")
    return out


def gn_stats(x_buf, c, off=0):
    lib = L.load()
    n, h, w, ld = x_buf.shape
    ws = th.empty(int(lib.gd_groupnorm_ws_floats(n, h * w, c)), dtype=th.float32, device=x_buf.device)
    st = th.empty((n, 32, 2), dtype=th.float32, device=x_buf.device)
    L.check(lib.gd_groupnorm_stats(C.c_void_p(x_buf.data_ptr() + 2 * off), ld, n, h * w, c, C.c_float(1e-5), vp(ws),
                                   vp(st), stream()), "gd_groupnorm_stats")
    return st


def gn_coef(st, gamma, beta, film, n, c):
    """Affine table [n][c/8][16] of GroupNorm32 (+FiLM) from finished statistics (gd_groupnorm_coef)."""
    coef = th.empty((n, c // 8, 16), dtype=th.float32, device=st.device)
    L.check(L.load().gd_groupnorm_coef(vp(st), vp(gamma), vp(beta), vp(film), film.shape[1] if film is not None else 0,
                                       n, c, vp(coef), stream()), "gd_groupnorm_coef")
    return coef


def gn_apply(x_buf, c, st, gamma, beta, *, film=None, silu=True, mode=L.GN_SAME, off=0, ld_out=None, out_off=0,
             aux=None):
    lib = L.load()
    n, h, w, ld = x_buf.shape
    ho, wo = (h // 2, w // 2) if mode == L.GN_AVGPOOL2 else (h * 2, w * 2) if mode == L.GN_UPSAMPLE2 else (h, w)
    ld_out = ld_out or c
    out = th.zeros((n, ho, wo, ld_out), dtype=th.float16, device=x_buf.device)
    L.check(lib.gd_groupnorm_apply(C.c_void_p(x_buf.data_ptr() + 2 * off), ld, vp(st), vp(gamma), vp(beta), vp(film),
                                   film.shape[1] if film is not None else 0,
                                   C.c_void_p(out.data_ptr() + 2 * out_off), ld_out, n, h, w, c, int(silu), mode,
                                   vp(aux), aux.shape[-1] if aux is not None else 0, stream()), "gd_groupnorm_apply")
    return out


def gn_bwd(x_buf, c, st, gamma, beta, dy_buf, *, film=None, silu=True, mode=L.GN_SAME, add_buf=None,
           add_mode=L.GN_SAME):
    lib = L.load()
    n, h, w, ld = x_buf.shape
    ws = th.empty(int(lib.gd_groupnorm_ws_floats(n, h * w, c)), dtype=th.float32, device=x_buf.device)
    dx = th.zeros((n, h, w, c), dtype=th.float16, device=x_buf.device)
    L.check(lib.gd_groupnorm_bwd(vp(x_buf), ld, vp(st), vp(gamma), vp(beta), vp(film),
                                 film.shape[1] if film is not None else 0, vp(dy_buf), dy_buf.shape[-1], vp(add_buf),
                                 add_buf.shape[-1] if add_buf is not None else 0, add_mode, vp(dx), c, vp(ws), n, h, w,
                                 c, int(silu), mode, stream()), "gd_groupnorm_bwd")
    return dx


def attention_fwd(qkv_buf, heads, order, want_lse=True):
    lib = L.load()
    n, t, ld = qkv_buf.shape
    out = th.zeros((n, t, heads * 64), dtype=th.float16, device=qkv_buf.device)
    lse = th.zeros((n, heads, t), dtype=th.float32, device=qkv_buf.device) if want_lse else None
    L.check(lib.gd_attention_fwd(vp(qkv_buf), ld, vp(out), heads * 64, vp(lse), n, t, heads, order, stream()),
            "gd_attention_fwd")
    return out, lse


def attention_bwd(qkv_buf, out, dout, lse, heads, order):
    lib = L.load()
    n, t, ld = qkv_buf.shape
    delta = th.zeros((n, heads, t), dtype=th.float32, device=qkv_buf.device)
    dqkv = th.zeros_like(qkv_buf)
    L.check(lib.gd_attention_bwd(vp(qkv_buf), ld, vp(out), out.shape[-1], vp(dout), dout.shape[-1], vp(lse), vp(delta),
                                 vp(dqkv), ld, n, t, heads, order, stream()), "gd_attention_bwd")
    return dqkv


def rel_err(a: th.Tensor, b: th.Tensor) -> float:
    """max-abs error relative to the reference's max-abs (the tolerance metric named in BASELINE.json)."""
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
