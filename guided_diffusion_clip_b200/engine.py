"""Host-side execution planner: turns a model spec (unet.py) into a static sequence of C-ABI kernel launches
over pre-allocated device buffers.  A plan is built once per (model, batch, resolution); running it is a
tight loop of ctypes calls on the current CUDA stream, which makes it CUDA-graph capturable as is.

Data layout in HBM:
  * activations: NHWC fp16 "channel views" (buffer, channel offset, C); the skip concatenations of
    unet.py:661 are physical buffers whose channel slices are written directly by their producers.
  * conv weights: packed once to fp16 [N_pad][K] (K = tap*C_in + c, then fused 1x1-skip channels).
  * everything fp32 in the reference (time/label embedding MLPs, emb_layers, GroupNorm parameters and
    statistics, the pool head, the diffusion update) stays fp32.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch as th

from . import _lib as L
from .unet import AttnSpec, ConvInSpec, ResampleSpec, ResSpec

GN_EPS = 1e-5


def norm_device(device) -> th.device:
    """Canonical device (cuda -> cuda:<current index>) so plan-cache keys agree between callers."""
    d = th.device(device)
    if d.type == "cuda" and d.index is None:
        d = th.device("cuda", th.cuda.current_device())
    return d


def _require_cuda(device) -> None:
    if th.device(device).type != "cuda":
        raise L.GdError(
            "guided_diffusion_clip_b200 only computes on CUDA (sm_100a); there is no CPU path. "
            f"Got device {device}.")


# ------------------------------------------------------------------------------------------------
# views and program
# ------------------------------------------------------------------------------------------------
@dataclass
class View:
    """NHWC fp16 channel view into a [N,H,W,LD] buffer."""
    buf: th.Tensor
    off: int
    c: int

    @property
    def n(self): return self.buf.shape[0]
    @property
    def h(self): return self.buf.shape[1]
    @property
    def w(self): return self.buf.shape[2]
    @property
    def ld(self): return self.buf.shape[3]
    @property
    def ptr(self): return self.buf.data_ptr() + 2 * self.off

    def slice(self, off: int, c: int) -> "View":
        assert off + c <= self.c
        return View(self.buf, self.off + off, c)

    def torch(self) -> th.Tensor:
        return self.buf[..., self.off:self.off + self.c]


def new_act(n, h, w, c, device) -> View:
    return View(th.empty((n, h, w, c), dtype=th.float16, device=device), 0, c)


class Program:
    """A recorded list of (C function, args) — replayed on the current stream of the plan's device."""

    def __init__(self, device=None):
        self.calls: List[Tuple[object, tuple, str]] = []
        self.keep: List[object] = []
        self.device = device
        self._measured: Optional[int] = None  # kernels launched by one run (counted by the library on the first run)

    def add(self, name: str, *args) -> None:
        fn = getattr(L.load(), name)
        self.calls.append((fn, args, name))

    def run(self) -> None:
        # the recorded launches embed raw pointers of the plan's device: always launch there, whatever device is
        # current in the caller (a launch on another GPU's stream would fault and poison the context)
        if self.device is not None and th.cuda.current_device() != self.device.index:
            with th.cuda.device(self.device):
                return self.run()
        stream = C.c_void_p(th.cuda.current_stream().cuda_stream)
        if os.environ.get("GD_B200_DEBUG_SYNC", "0") == "1":  # locate a faulting launch: sync after every call
            for i, (fn, args, name) in enumerate(self.calls):
                rc = fn(*args, stream)
                if rc != 0:
                    L.check(rc, name)
                try:
                    th.cuda.synchronize()
                except Exception as e:  # noqa: BLE001
                    raise L.GdError(f"launch #{i} {name} faulted: {e}") from e
            return
        before = int(L.load().gd_launch_count()) if self._measured is None else 0
        for fn, args, name in self.calls:
            rc = fn(*args, stream)
            if rc != 0:
                L.check(rc, name)
        if self._measured is None:
            self._measured = int(L.load().gd_launch_count()) - before

    @property
    def launches(self) -> int:
        if self._measured is not None:  # exact (a split-K conv is two launches, decided at launch time)
            return self._measured
        per = {"gd_groupnorm_stats": 2, "gd_groupnorm_bwd": 3, "gd_attention_bwd": 3, "gd_attnpool_fwd": 6,
               "gd_attnpool_bwd": 4}
        return sum(per.get(name, 1) for _, _, name in self.calls)


def _p(t: Optional[th.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else None


# ------------------------------------------------------------------------------------------------
# weight packing (host side, once per plan)
# ------------------------------------------------------------------------------------------------
def _pad_rows(w2d: th.Tensor) -> th.Tensor:
    n = w2d.shape[0]
    n_pad = (n + 15) // 16 * 16
    if n_pad != n:
        w2d = th.cat([w2d, w2d.new_zeros(n_pad - n, w2d.shape[1])], 0)
    return w2d.to(th.float16).contiguous()


def pack_conv3x3(w: th.Tensor, skip_w: Optional[th.Tensor] = None) -> th.Tensor:
    """OIHW [Co,Ci,3,3] -> [Co_pad][9*Ci (+ Cskip)] with k = (ky*3+kx)*Ci + ci; A-side tap offset (ky-1,kx-1)."""
    co, ci = w.shape[:2]
    m = w.float().permute(0, 2, 3, 1).reshape(co, 9 * ci)
    if skip_w is not None:
        m = th.cat([m, skip_w.float().reshape(co, -1)], 1)
    return _pad_rows(m)


def pack_conv3x3_bwd(w: th.Tensor) -> th.Tensor:
    """Weights of the data-gradient conv: dX[ci] = sum_{tap',co} dY[y+dy',x+dx'][co] * W[co][ci][2-ky'][2-kx']."""
    co, ci = w.shape[:2]
    m = w.float().flip(2, 3).permute(1, 2, 3, 0).reshape(ci, 9 * co)
    return _pad_rows(m)


def pack_tap_expand(w: th.Tensor) -> th.Tensor:
    """OIHW [Co,Ci,3,3] -> [64][Ci] fp16, row = (ky*3+kx)*Co + co (zero rows above 9*Co): the 3x3 conv as one 1x1 GEMM
    whose 9*Co output columns are gathered by gd_tap_gather3x3 (narrow-output convs only, 9*Co <= 63)."""
    co, ci = w.shape[:2]
    m = w.float().permute(2, 3, 0, 1).reshape(9 * co, ci)
    return th.cat([m, m.new_zeros(64 - 9 * co, ci)], 0).to(th.float16).contiguous()


NARROW_COUT = 7  # 9 * cout <= 63: the tap-expanded GEMM still fits one 64-column tile


def pack_conv_in(w: th.Tensor) -> th.Tensor:
    """First-layer weights [Co,Ci,3,3] (Ci*9 <= 64) -> [Co_pad][64] matching gd_im2col3x3_small_cin: k = (ky*3+kx)*Ci + ci."""
    co, ci = w.shape[:2]
    m = w.float().permute(0, 2, 3, 1).reshape(co, 9 * ci)
    m = th.cat([m, m.new_zeros(co, 64 - 9 * ci)], 1)
    return _pad_rows(m)


def pack_1x1(w: th.Tensor) -> th.Tensor:
    return _pad_rows(w.float().reshape(w.shape[0], -1))


def pack_1x1_bwd(w: th.Tensor) -> th.Tensor:
    return _pad_rows(w.float().reshape(w.shape[0], -1).t())


class Emitter:
    """Shared emission helpers for the UNet and classifier plans."""

    def __init__(self, model, n: int, device):
        _require_cuda(device)
        L.load()
        self.model = model
        self.n = n
        self.device = device = norm_device(device)
        self.P: Dict[str, th.Tensor] = {k: v.detach() for k, v in model.named_parameters()}
        for k, v in self.P.items():
            if norm_device(v.device) != device:
                # same situations in which the reference raises a device-mismatch RuntimeError; here the kernels would
                # otherwise receive a host / foreign-GPU pointer
                raise L.GdError(f"parameter {k} lives on {v.device} but the input is on {device}: move the model with "
                                ".to(device) before calling it")
        self.prog = Program(device)
        self._scratch: Dict[str, th.Tensor] = {}
        self._f32: Dict[str, th.Tensor] = {}
        self.keep: List[th.Tensor] = []
        self.gn_ws = None
        self._gn_ws_floats = 0
        self._producers: Dict[tuple, tuple] = {}   # (buffer ptr, channel offset, C) -> (ConvDesc, n, h, w)
        self._stats_bufs: Dict[int, th.Tensor] = {}  # id(ConvDesc) -> fused GroupNorm partials of that conv
        self._splitk_ws: Optional[th.Tensor] = None  # split-K workspace shared by the plan's (serialised) convs

    # ---- buffers ------------------------------------------------------------------------------
    def scratch(self, role: str, n, h, w, c) -> View:
        """Reusable fp16 scratch, grown to the largest request per role (allocated in finalize())."""
        need = n * h * w * c
        cur = self._scratch.get(role)
        if cur is None or cur.numel() < need:
            self._scratch[role] = th.empty(need, dtype=th.float16, device=self.device)
            # earlier views of this role keep their (smaller) tensor alive through Program.keep
        buf = self._scratch[role][:need].view(n, h, w, c)
        self.keep.append(buf)
        return View(buf, 0, c)

    def act(self, n, h, w, c) -> View:
        """A dedicated activation buffer OWNED by this emitter (kernels only see raw pointers, so every buffer a
        recorded launch touches must stay referenced for the lifetime of the plan)."""
        v = new_act(n, h, w, c, self.device)
        self.keep.append(v.buf)
        return v

    def f32(self, name: str) -> th.Tensor:
        """fp32 copy of a parameter (biases / GN affine / linear weights); fp16-rounded if the model holds it in fp16."""
        t = self._f32.get(name)
        if t is None:
            t = self.P[name].float().contiguous()
            self._f32[name] = t
        return t

    def stats_buf(self) -> th.Tensor:
        t = th.empty((self.n, 32, 2), dtype=th.float32, device=self.device)
        self.keep.append(t)
        return t

    def _gn_workspace(self) -> th.Tensor:
        if self.gn_ws is None:
            floats = int(L.load().gd_groupnorm_ws_floats(self.n, 1, 32))
            self.gn_ws = th.empty(floats, dtype=th.float32, device=self.device)
        return self.gn_ws

    # ---- op emitters --------------------------------------------------------------------------
    def conv(self, a0: View, wpack: th.Tensor, bias: Optional[th.Tensor], cout: int, out, *, taps=9,
             a1: Optional[View] = None, res: Optional[View] = None, res_mode=L.RES_NONE, out_mode=L.OUT_NHWC_F16,
             out_scale=1.0, geom: Optional[Tuple[int, int, int]] = None, gn: Optional[dict] = None) -> None:
        """gn = dict(mode, coef, silu): a0 is the RAW tensor and GroupNorm (+FiLM, +SiLU, + nearest x2 for mode
        CONV_GN_UPSAMPLE2) is applied on its way into the tensor cores (gd_conv_desc.gn_*; coef from gn_stats)."""
        n, h, w = geom if geom is not None else (a0.n, a0.h, a0.w)
        d = L.ConvDesc()
        d.a0, d.c0, d.ld0, d.taps = a0.ptr, a0.c, a0.ld, taps
        if a1 is not None:
            d.a1, d.c1, d.ld1 = a1.ptr, a1.c, a1.ld
        else:
            d.a1, d.c1, d.ld1 = None, 0, 0
        d.n, d.h, d.w = n, h, w
        d.wpack, d.k_total, d.n_pad = wpack.data_ptr(), wpack.shape[1], wpack.shape[0]
        assert wpack.shape[1] == taps * a0.c + (a1.c if a1 is not None else 0), (wpack.shape, taps, a0.c)
        d.bias = bias.data_ptr() if bias is not None else None
        d.cout = cout
        if res is not None:
            d.res, d.ld_res, d.res_mode = res.ptr, res.ld, res_mode
        else:
            d.res, d.ld_res, d.res_mode = None, 0, L.RES_NONE
        if out_mode == L.OUT_NHWC_F16:
            d.out, d.ld_out = out.ptr, out.ld
        else:
            d.out, d.ld_out = out.data_ptr(), 0
        d.out_mode = out_mode
        d.bn = 0
        d.out_scale = out_scale
        d.stats_out = None
        if gn is not None:
            d.gn_mode, d.gn_silu = gn["mode"], int(gn.get("silu", True))
            d.gn_coef = gn["coef"].data_ptr()
            self.keep.append(gn["coef"])
        need = int(L.load().gd_conv_splitk_ws_bytes(C.byref(d))) if os.environ.get("GD_B200_SPLITK", "0") == "1" else 0
        if need > 0:
            # few pixel tiles, long K (8x8 / 16x16 layers at small batch): lend the launch a workspace to split K over
            # the idle SMs.  One workspace per plan: its launches are serialised on one stream.  Opt-in
            # (GD_B200_SPLITK=1): the split count follows the batch, which costs bit-for-bit batch invariance.
            if self._splitk_ws is None or self._splitk_ws.numel() * 4 < need:
                self._splitk_ws = th.empty((need + 3) // 4, dtype=th.float32, device=self.device)
            self.keep.append(self._splitk_ws)
            d.splitk_ws, d.splitk_ws_bytes = self._splitk_ws.data_ptr(), self._splitk_ws.numel() * 4
        self.keep += [wpack, bias, d]
        self.prog.add("gd_conv_igemm", C.byref(d))
        self.last_scale_slot = ("desc", d, None)
        if out_mode == L.OUT_NHWC_F16 and cout == out.c and cout % 64 == 0 and wpack.shape[0] == cout:
            # latest writer of this channel view: a later GroupNorm over it can ask this conv for fused statistics
            self._producers[(out.buf.data_ptr(), out.off, out.c)] = (d, n, h, w)

    def conv3x3_narrow(self, a0: View, w_oihw: th.Tensor, bias: Optional[th.Tensor], out: th.Tensor, *,
                       out_scale=1.0, geom: Optional[Tuple[int, int, int]] = None) -> None:
        """3x3 conv to fp32 NCHW with cout <= NARROW_COUT as a tap-expanded 1x1 GEMM (64 columns, the fast fp16 NHWC
        TMA-store epilogue) + tap gather: a tensor-core tile with 3-6 useful columns of 16 runs at 35-55 TFLOP/s."""
        n, h, w = geom if geom is not None else (a0.n, a0.h, a0.w)
        cout = int(w_oihw.shape[0])
        ytap = new_act(n, h, w, 64, self.device)
        self.keep.append(ytap.buf)
        self.conv(a0, pack_tap_expand(w_oihw), None, 64, ytap, taps=1, geom=geom)
        self.keep.append(bias)
        self.prog.add("gd_tap_gather3x3", C.c_void_p(ytap.ptr), ytap.ld, _p(bias), _p(out), n, cout, h, w,
                      C.c_float(float(out_scale)))
        self.last_scale_slot = ("arg", len(self.prog.calls) - 1, 8)

    def _producers_of(self, x: View):
        """The conv launch(es) that wrote view x: one conv, or two convs writing adjacent channel slices (skip concat)."""
        base = x.buf.data_ptr()
        hit = self._producers.get((base, x.off, x.c))
        if hit is not None:
            return [(hit, x.c)]
        for (b, off, c), prod in self._producers.items():
            if b == base and off == x.off and c < x.c:
                rest = self._producers.get((base, x.off + c, x.c - c))
                if rest is not None:
                    return [(prod, c), (rest, x.c - c)]
        return None

    def gn_stats(self, x: View, stats: th.Tensor, affine: Optional[dict] = None) -> Optional[th.Tensor]:
        """GroupNorm32 statistics of x.  Preferred: the producing conv(s) emit per-row-block partial sums from their
        epilogue (no extra pass over the tensor) and a tiny finalize kernel turns them into mean/rstd; otherwise the
        two-launch statistics kernels read the tensor once.
        affine = dict(gamma, beta, film, film_ld): additionally produce (and return) the per-channel affine table
        [n][C/8][16] a conv with a fused GroupNorm operand consumes — written by the finalize kernel itself where it
        runs, by gd_groupnorm_coef otherwise."""
        lib = L.load()
        coef = None
        aff = (None, None, None, 0, None)
        if affine is not None:
            coef = th.empty((x.n, x.c // 8, 16), dtype=th.float32, device=self.device)
            self.keep += [coef, affine["gamma"], affine["beta"]]
            film = affine.get("film")
            aff = (_p(affine["gamma"]), _p(affine["beta"]), C.c_void_p(film) if film else None,
                   affine.get("film_ld", 0) if film else 0, _p(coef))
        parts = self._producers_of(x) if os.environ.get("GD_B200_NO_FUSED_STATS", "0") != "1" else None
        rpi = C.c_int32(0)
        rows = int(lib.gd_conv_stats_rows(x.n, x.h, x.w, C.byref(rpi))) if parts else 0
        if parts and rows > 0 and (x.c // 32) % 4 == 0 and all((n_, h_, w_) == (x.n, x.h, x.w)
                                                                for (_, n_, h_, w_), _c in parts):
            bufs = []
            for (desc, _n, _h, _w), c in parts:
                buf = self._stats_bufs.get(id(desc))
                if buf is None:
                    buf = th.zeros((rows, c // 4, 2), dtype=th.float32, device=self.device)
                    self._stats_bufs[id(desc)] = buf
                    desc.stats_out = buf.data_ptr()
                bufs.append((buf, c))
            (p0, c0), (p1, c1) = bufs[0], (bufs[1] if len(bufs) > 1 else (None, 0))
            self.prog.add("gd_groupnorm_finalize_partials", _p(p0), c0, c0 // 4, _p(p1), c1, c1 // 4, rpi.value, x.n,
                          x.h * x.w, C.c_float(GN_EPS), _p(stats), *aff)
            return coef
        self.prog.add("gd_groupnorm_stats", C.c_void_p(x.ptr), x.ld, x.n, x.h * x.w, x.c, C.c_float(GN_EPS),
                      _p(self._gn_workspace()), _p(stats))
        if coef is not None:
            self.prog.add("gd_groupnorm_coef", _p(stats), aff[0], aff[1], aff[2], aff[3], x.n, x.c, aff[4])
        return coef

    def gn_apply(self, x: View, stats, gamma, beta, out: View, *, silu: bool, film=None, film_ld=0,
                 mode=L.GN_SAME, aux: Optional[View] = None) -> None:
        self.keep += [gamma, beta]
        self.prog.add("gd_groupnorm_apply", C.c_void_p(x.ptr), x.ld, _p(stats), _p(gamma), _p(beta),
                      C.c_void_p(film) if film else None, film_ld, C.c_void_p(out.ptr), out.ld, x.n, x.h, x.w, x.c,
                      int(silu), mode, C.c_void_p(aux.ptr) if aux is not None else None,
                      aux.ld if aux is not None else 0)

    def gn_bwd(self, x: View, stats, gamma, beta, dy: View, dx: View, *, silu: bool, film=None, film_ld=0,
               mode=L.GN_SAME, add: Optional[View] = None, add_mode=L.GN_SAME) -> None:
        self.prog.add("gd_groupnorm_bwd", C.c_void_p(x.ptr), x.ld, _p(stats), _p(gamma), _p(beta),
                      C.c_void_p(film) if film else None, film_ld, C.c_void_p(dy.ptr), dy.ld,
                      C.c_void_p(add.ptr) if add is not None else None, add.ld if add is not None else 0, add_mode,
                      C.c_void_p(dx.ptr), dx.ld, _p(self._gn_workspace()), x.n, x.h, x.w, x.c, int(silu), mode)

    def linear(self, x: th.Tensor, w: th.Tensor, b, y: th.Tensor, *, m, k, n, ldx=None, ldy=None, add=None,
               silu_in=False, silu_out=False) -> None:
        self.keep += [x, w, b, y, add]
        self.prog.add("gd_linear_f32", _p(x), ldx or k, _p(w), _p(b), _p(add), (add.shape[-1] if add is not None else 0),
                      _p(y), ldy or n, m, k, n, int(silu_in), int(silu_out))

    # ---- embeddings ---------------------------------------------------------------------------
    def emit_embedding(self, t_buf: th.Tensor, cond: Optional[th.Tensor], spec) -> th.Tensor:
        """timestep_embedding -> time_embed MLP (+ label embedding) -> SiLU -> ONE batched emb_layers GEMM
        producing every ResBlock's (scale, shift).  Returns film_all [n, film_total] fp32."""
        n, mc, e = self.n, spec.model_channels, spec.emb_dim
        dev = self.device
        temb = th.empty((n, mc), dtype=th.float32, device=dev)
        h1 = th.empty((n, e), dtype=th.float32, device=dev)
        emb_silu = th.empty((n, e), dtype=th.float32, device=dev)
        self.prog.add("gd_timestep_embedding", _p(t_buf), _p(temb), n, mc)
        self.keep += [t_buf, temb]
        self.linear(temb, self.f32("time_embed.0.weight"), self.f32("time_embed.0.bias"), h1, m=n, k=mc, n=e,
                    silu_out=True)
        lab = None
        model = self.model
        if getattr(model, "num_classes", None) is not None:
            lab = th.empty((n, e), dtype=th.float32, device=dev)
            if model.label_mlp:
                nc = model.num_classes
                l1 = th.empty((n, e), dtype=th.float32, device=dev)
                self.linear(cond, self.f32("label_emb.0.weight"), self.f32("label_emb.0.bias"), l1, m=n, k=nc, n=e,
                            silu_out=True)
                self.linear(l1, self.f32("label_emb.2.weight"), self.f32("label_emb.2.bias"), lab, m=n, k=e, n=e)
            else:
                table = self.f32("label_emb.weight")
                self.keep += [table, cond, lab]
                self.prog.add("gd_embedding_gather", _p(table), _p(cond), _p(lab), n, e, table.shape[0])
        # emb = time_embed(...) + label ; every consumer applies SiLU first (unet.py:200) -> store SiLU(emb)
        self.linear(h1, self.f32("time_embed.2.weight"), self.f32("time_embed.2.bias"), emb_silu, m=n, k=e, n=e,
                    add=lab, silu_out=True)
        blocks = spec.res_blocks()  # rows of 2*C (scale, shift) with FiLM, of C (additive embedding) without
        w_all = th.cat([self.f32(f"{r.key}.emb_layers.1.weight") for r in blocks], 0).contiguous()
        b_all = th.cat([self.f32(f"{r.key}.emb_layers.1.bias") for r in blocks], 0).contiguous()
        film_all = th.empty((n, spec.film_total), dtype=th.float32, device=dev)
        self.linear(emb_silu, w_all, b_all, film_all, m=n, k=e, n=spec.film_total)
        return film_all

    # ---- blocks -------------------------------------------------------------------------------
    def conv_in(self, l: ConvInSpec, x_nchw: th.Tensor, out: View) -> None:
        """First layer (unet.py:483).  Preferred: gd_conv_in3x3, one launch from the fp32 NCHW input with register
        accumulators (the layer is bound by writing its output).  Shapes it does not cover (C_out not a multiple of 64,
        fewer than 128 pixels per image): im2col to one 64-wide K block, then a K=64 GEMM on the tcgen05 kernel."""
        self.keep.append(x_nchw)
        wp, bias = pack_conv_in(self.P[f"{l.key}.weight"]), self.f32(f"{l.key}.bias")
        if (l.cout % 64 == 0 and (out.h * out.w) % 128 == 0 and out.ld % 8 == 0 and out.off % 8 == 0
                and os.environ.get("GD_B200_NO_CONV_IN", "0") != "1"):
            d = L.ConvInDesc()
            d.x, d.wpack, d.bias, d.out, d.stats_out = x_nchw.data_ptr(), wp.data_ptr(), bias.data_ptr(), out.ptr, None
            d.n, d.cin, d.h, d.w, d.cout, d.ld_out = out.n, l.cin, out.h, out.w, l.cout, out.ld
            self.keep += [wp, bias, d]
            self.prog.add("gd_conv_in3x3", C.byref(d))
            if l.cout == out.c:
                self._producers[(out.buf.data_ptr(), out.off, out.c)] = (d, out.n, out.h, out.w)
            return
        cols = self.scratch("im2col", out.n, out.h, out.w, 64)
        self.prog.add("gd_im2col3x3_small_cin", _p(x_nchw), C.c_void_p(cols.ptr), cols.ld, out.n, l.cin, out.h, out.w)
        self.conv(cols, wp, bias, l.cout, out, taps=1)

    def res_block(self, r: ResSpec, x: View, out: View, film_all: th.Tensor, tape: Optional[list] = None) -> None:
        """ResBlock._forward (unet.py:236-256) in 6 launches: GN stats, GN-apply(+SiLU, +pool/upsample), conv,
        GN stats, GN-apply(+FiLM+SiLU), conv (+1x1 skip as extra K blocks | identity residual in the epilogue)."""
        n = self.n
        if r.mode == "down":
            ho, wo, gmode, rmode = x.h // 2, x.w // 2, L.GN_AVGPOOL2, L.RES_AVGPOOL2
        elif r.mode == "up":
            ho, wo, gmode, rmode = x.h * 2, x.w * 2, L.GN_UPSAMPLE2, L.RES_UPSAMPLE2
        else:
            ho, wo, gmode, rmode = x.h, x.w, L.GN_SAME, L.RES_SAME
        assert (out.h, out.w, out.c) == (ho, wo, r.cout), (out.buf.shape, ho, wo, r.cout)
        k = r.key
        st1, st2 = self.stats_buf(), self.stats_buf()
        # GroupNorm -> SiLU (-> nearest x2) -> conv: the normalisation runs inside the conv's operand path wherever the
        # conv tiles whole 16 x 8 patches (every resolution >= 16 x 16); the normalised tensor then never exists in HBM.
        # Down blocks average 4 activated pixels per conv input and keep the separate pass (which also emits
        # avgpool(x), the block's identity residual, as a side output).
        fuse = (os.environ.get("GD_B200_NO_GN_FUSE", "0") != "1" and bool(L.load().gd_conv_gn_fusable(ho, wo))
                and x.ld % 8 == 0 and x.off % 8 == 0)
        g1, b1 = self.f32(f"{k}.in_layers.0.weight"), self.f32(f"{k}.in_layers.0.bias")
        fuse1 = fuse and r.mode != "down"
        coef1 = self.gn_stats(x, st1, affine=dict(gamma=g1, beta=b1) if fuse1 else None)
        h1 = self.act(n, ho, wo, r.cout) if tape is not None else self.scratch("h1", n, ho, wo, r.cout)
        w1 = pack_conv3x3(self.P[f"{k}.in_layers.2.weight"])
        x_pool = None
        if fuse1:
            self.conv(x, w1, self.f32(f"{k}.in_layers.2.bias"), r.cout, h1, geom=(n, ho, wo),
                      gn=dict(mode=L.CONV_GN_UPSAMPLE2 if r.mode == "up" else L.CONV_GN_SAME, coef=coef1, silu=True))
        else:
            a = self.scratch("gn_out", n, ho, wo, r.cin)
            x_pool = self.scratch("x_pool", n, ho, wo, r.cin) if r.mode == "down" else None
            self.gn_apply(x, st1, g1, b1, a, silu=True, mode=gmode, aux=x_pool)
            self.conv(a, w1, self.f32(f"{k}.in_layers.2.bias"), r.cout, h1)
        film_ptr = film_all.data_ptr() + 4 * r.film_offset
        film = self.model.use_scale_shift_norm
        if not film:
            # h = h + emb_out before the second GroupNorm (unet.py:253-255): one in-place pass; the statistics the
            # conv fused into its epilogue describe the tensor before the add, so they are recomputed
            self.prog.add("gd_add_emb_nhwc", C.c_void_p(h1.ptr), h1.ld, C.c_void_p(film_ptr), film_all.shape[1], n,
                          ho * wo, r.cout)
            self._producers.pop((h1.buf.data_ptr(), h1.off, h1.c), None)
        g2, b2 = self.f32(f"{k}.out_layers.0.weight"), self.f32(f"{k}.out_layers.0.bias")
        coef2 = self.gn_stats(h1, st2, affine=dict(gamma=g2, beta=b2, film=film_ptr if film else None,
                                                   film_ld=film_all.shape[1]) if fuse else None)
        if fuse:
            b, gn2 = h1, dict(mode=L.CONV_GN_SAME, coef=coef2, silu=True)
            self.keep.append(film_all)
        else:
            b, gn2 = self.scratch("gn_out2", n, ho, wo, r.cout), None
            self.gn_apply(h1, st2, g2, b2, b, silu=True, film=film_ptr if film else None,
                          film_ld=film_all.shape[1] if film else 0)
        if r.has_skip_conv:
            assert r.mode == "none"
            w2 = pack_conv3x3(self.P[f"{k}.out_layers.3.weight"], self.P[f"{k}.skip_connection.weight"])
            bias2 = (self.f32(f"{k}.out_layers.3.bias") + self.f32(f"{k}.skip_connection.bias")).contiguous()
            self.conv(b, w2, bias2, r.cout, out, a1=x, gn=gn2)
        else:
            w2 = pack_conv3x3(self.P[f"{k}.out_layers.3.weight"])
            if x_pool is not None:
                self.conv(b, w2, self.f32(f"{k}.out_layers.3.bias"), r.cout, out, res=x_pool, res_mode=L.RES_SAME,
                          gn=gn2)
            elif r.mode == "down":
                raise AssertionError("down block without the pooled residual")
            else:
                self.conv(b, w2, self.f32(f"{k}.out_layers.3.bias"), r.cout, out, res=x, res_mode=rmode, gn=gn2)
        if tape is not None:
            # without FiLM the saved h1 already holds h + emb_out (what the second GroupNorm saw): d(h + e)/dh = 1
            tape.append(("res", r, x, st1, h1, st2, film_ptr if film else None, film_all.shape[1] if film else 0, out))

    def resample(self, l: ResampleSpec, x: View, out: View) -> None:
        """Downsample.op (3x3 stride-2 conv, unet.py:125-136) as a strided gather + taps=1 GEMM over K = 9*C;
        Upsample (unet.py:100-110) as a nearest x2 copy + the regular 3x3 conv."""
        n, c = x.n, l.ch
        wp = pack_conv3x3(self.P[f"{l.key}.weight"])
        bias = self.f32(f"{l.key}.bias")
        if l.mode == "down":
            ho, wo = (x.h - 1) // 2 + 1, (x.w - 1) // 2 + 1
            cols = self.scratch("im2col_s2", n, ho, wo, 9 * c)
            self.prog.add("gd_im2col3x3_s2_nhwc", C.c_void_p(x.ptr), x.ld, C.c_void_p(cols.ptr), cols.ld, n, x.h, x.w, c)
            self.conv(cols, wp, bias, c, out, taps=1)
        else:
            up = self.scratch("up2", n, 2 * x.h, 2 * x.w, c)
            self.prog.add("gd_upsample2_nhwc", C.c_void_p(x.ptr), x.ld, C.c_void_p(up.ptr), up.ld, n, x.h, x.w, c)
            self.conv(up, wp, bias, c, out)

    def attn_block(self, a: AttnSpec, x: View, out: View, tape: Optional[list] = None) -> None:
        """AttentionBlock._forward (unet.py:299-305): GN, qkv 1x1, fused attention, proj 1x1 + residual."""
        hd = a.ch // a.heads
        if a.ch != a.heads * hd or hd % 16 or hd > 256 or (hd > 128 and hd % 32):
            raise NotImplementedError(f"attention head dim {a.ch}/{a.heads}: CUDA kernels exist for 16..128 in steps of "
                                      "16 and 160 / 192 / 224 / 256")
        if hd != 64 and tape is not None:
            raise NotImplementedError(f"attention data-gradient exists for 64-wide heads only (got {hd}; the "
                                      "classifier factory hard-codes 64, script_util.py:265)")
        n, h, w = x.n, x.h, x.w
        k = a.key
        st = self.stats_buf()
        g = self.scratch("gn_out", n, h, w, a.ch)
        self.gn_stats(x, st)
        self.gn_apply(x, st, self.f32(f"{k}.norm.weight"), self.f32(f"{k}.norm.bias"), g, silu=False)
        keep = tape is not None
        qkv = self.act(n, h, w, 3 * a.ch) if keep else self.scratch("qkv", n, h, w, 3 * a.ch)
        att = self.act(n, h, w, a.ch) if keep else self.scratch("att", n, h, w, a.ch)
        lse = th.empty((n, a.heads, h * w), dtype=th.float32, device=self.device) if keep else None
        self.conv(g, pack_1x1(self.P[f"{k}.qkv.weight"]), self.f32(f"{k}.qkv.bias"), 3 * a.ch, qkv, taps=1)
        order = L.QKV_NEW if a.new_order else L.QKV_LEGACY
        self.keep.append(lse)
        if hd == 64 and ((h * w) % 64 == 0 or keep):  # (token counts off the 64 grid: the any-length kernel, forward only)
            self.prog.add("gd_attention_fwd", C.c_void_p(qkv.ptr), qkv.ld, C.c_void_p(att.ptr), att.ld, _p(lse), n, h * w,
                          a.heads, order)
        else:
            self.prog.add("gd_attention_fwd_hd", C.c_void_p(qkv.ptr), qkv.ld, C.c_void_p(att.ptr), att.ld, _p(lse), n,
                          h * w, a.heads, hd, order)
        self.conv(att, pack_1x1(self.P[f"{k}.proj_out.weight"]), self.f32(f"{k}.proj_out.bias"), a.ch, out, taps=1,
                  res=x, res_mode=L.RES_SAME)
        if tape is not None:
            tape.append(("attn", a, x, st, qkv, att, lse, out))


# ------------------------------------------------------------------------------------------------
# UNet forward plan
# ------------------------------------------------------------------------------------------------
class UNetPlan:
    """UNetModel.forward (unet.py:635-664) for a fixed (batch, H, W)."""

    def __init__(self, model, n: int, h: int, w: int, device):
        em = Emitter(model, n, device)
        self.em = em
        spec = model.spec
        dev = device
        self.x_in = th.empty((n, model.in_channels, h, w), dtype=th.float32, device=dev)
        self.t_in = th.empty((n,), dtype=th.float32, device=dev)
        self.cond_in = None
        if model.num_classes is not None:
            self.cond_in = (th.empty((n, model.num_classes), dtype=th.float32, device=dev) if model.label_mlp
                            else th.empty((n,), dtype=th.int64, device=dev))
        self.out = th.empty((n, model.out_channels, h, w), dtype=th.float32, device=dev)
        film_all = em.emit_embedding(self.t_in, self.cond_in, spec)

        # --- geometry of every input-block output (the hs stack, unet.py:648,658) -------------------
        hs_shapes = []  # (c, h, w)
        ch, hh, ww = None, h, w
        for blk in spec.input_blocks:
            for l in blk:
                if isinstance(l, ConvInSpec):
                    ch = l.cout
                elif isinstance(l, ResSpec):
                    ch = l.cout
                    if l.mode == "down":
                        hh, ww = hh // 2, ww // 2
                elif isinstance(l, ResampleSpec):
                    hh, ww = (hh - 1) // 2 + 1, (ww - 1) // 2 + 1
            hs_shapes.append((ch, hh, ww))
        # --- concat buffers: output block j consumes cat([h, hs.pop()]) ------------------------------
        n_out = len(spec.output_blocks)
        cat_bufs: List[View] = []
        hcur = (ch, hh, ww)  # middle block output geometry
        hs_stack = list(hs_shapes)
        for j, blk in enumerate(spec.output_blocks):
            sc, sh, sw = hs_stack.pop()
            assert (sh, sw) == (hcur[1], hcur[2]), "skip / h resolution mismatch"
            cat_bufs.append(em.act(n, sh, sw, hcur[0] + sc))
            c2, h2, w2 = hcur
            for l in blk:
                if isinstance(l, ResSpec):
                    c2 = l.cout
                    if l.mode == "up":
                        h2, w2 = h2 * 2, w2 * 2
                elif isinstance(l, ResampleSpec):
                    h2, w2 = h2 * 2, w2 * 2
            hcur = (c2, h2, w2)
        final_view = em.act(n, hcur[1], hcur[2], hcur[0])

        def hs_view(i: int) -> View:
            # hs[i] is popped by output block j = n_out-1-i and sits after h's channels
            j = n_out - 1 - i
            cb = cat_bufs[j]
            c_skip = hs_shapes[i][0]
            return cb.slice(cb.c - c_skip, c_skip)

        def run_block(layers, x: View, dst: View):
            cur = x
            for li, l in enumerate(layers):
                last = li == len(layers) - 1
                if isinstance(l, ResSpec):
                    oh, ow = ((cur.h // 2, cur.w // 2) if l.mode == "down" else
                              (cur.h * 2, cur.w * 2) if l.mode == "up" else (cur.h, cur.w))
                    o = dst if last else em.scratch(f"blk{li % 2}", n, oh, ow, l.cout)
                    em.res_block(l, cur, o, film_all)
                elif isinstance(l, ResampleSpec):
                    oh, ow = ((cur.h - 1) // 2 + 1, (cur.w - 1) // 2 + 1) if l.mode == "down" else (cur.h * 2, cur.w * 2)
                    o = dst if last else em.scratch(f"blk{li % 2}", n, oh, ow, l.ch)
                    em.resample(l, cur, o)
                else:
                    o = dst if last else em.scratch(f"blk{li % 2}", n, cur.h, cur.w, l.ch)
                    em.attn_block(l, cur, o)
                cur = o
            return cur

        # --- input blocks ---------------------------------------------------------------------------
        cur: Optional[View] = None
        for i, blk in enumerate(spec.input_blocks):
            dst = hs_view(i)
            if isinstance(blk[0], ConvInSpec):
                em.conv_in(blk[0], self.x_in, dst)
                cur = dst
            else:
                cur = run_block(blk, cur, dst)
        # --- middle ---------------------------------------------------------------------------------
        mid_dst = cat_bufs[0].slice(0, cat_bufs[0].c - hs_shapes[-1][0]) if n_out else final_view
        cur = run_block(spec.middle_block, cur, mid_dst)
        # --- output blocks --------------------------------------------------------------------------
        for j, blk in enumerate(spec.output_blocks):
            if j + 1 < n_out:
                nxt = cat_bufs[j + 1]
                dst = nxt.slice(0, nxt.c - hs_shapes[n_out - 2 - j][0])
            else:
                dst = final_view
            cur = run_block(blk, cat_bufs[j], dst)
        # --- out head: GN -> SiLU -> conv3x3 (fp32 in the reference, unet.py:613-617,663-664) ---------
        st = em.stats_buf()
        g = em.scratch("gn_out", n, cur.h, cur.w, cur.c)
        em.gn_stats(cur, st)
        em.gn_apply(cur, st, em.f32("out.0.weight"), em.f32("out.0.bias"), g, silu=True)
        if model.out_channels <= NARROW_COUT:
            em.conv3x3_narrow(g, em.P["out.2.weight"], em.f32("out.2.bias"), self.out)
        else:
            em.conv(g, pack_conv3x3(em.P["out.2.weight"]), em.f32("out.2.bias"), model.out_channels, self.out,
                    out_mode=L.OUT_NCHW_F32)
        self.prog = em.prog

    def load_inputs(self, x, timesteps, cond) -> None:
        self.x_in.copy_(x)
        self.t_in.copy_(timesteps)
        if self.cond_in is not None:
            self.cond_in.copy_(cond)

    def run(self, x, timesteps, cond) -> th.Tensor:
        self.load_inputs(x, timesteps, cond)
        self.prog.run()
        return self.out


def bilinear_concat(x: th.Tensor, low_res: th.Tensor) -> th.Tensor:
    """cat([x, F.interpolate(low_res, (H,W), mode='bilinear')], 1) (unet.py:677-680) via the CUDA kernel."""
    _require_cuda(x.device)
    if low_res.device != x.device:
        raise L.GdError(f"low_res lives on {low_res.device} but x on {x.device} (load_data_for_worker yields CPU tensors: "
                        "move them with .to(device) like scripts/super_res_sample.py:46)")
    n, c, h, w = x.shape
    cl = low_res.shape[1]
    out = th.empty((n, c + cl, h, w), dtype=th.float32, device=x.device)
    out[:, :c].copy_(x)
    lr = low_res.float().contiguous()
    with th.cuda.device(x.device):
        stream = C.c_void_p(th.cuda.current_stream().cuda_stream)
        L.check(L.load().gd_bilinear_upsample_nchw(_p(lr), _p(out), n, cl, lr.shape[2], lr.shape[3], h, w, c + cl, c,
                                                   stream), "gd_bilinear_upsample_nchw")
    return out


# ------------------------------------------------------------------------------------------------
# classifier forward + dX backward plan
# ------------------------------------------------------------------------------------------------
class ClassifierPlan:
    """EncoderUNetModel.forward (unet.py:872-895) with every tensor the data-gradient needs kept resident
    (no recompute, no parameter gradients — unlike the reference's CheckpointFunction, nn.py:152-170), and
    the matching backward program producing d(sum selected log-probs)/dx in fp32 NCHW."""

    LOSS_SCALE = 256.0  # static fp16 gradient scale, undone in the last conv's epilogue

    def __init__(self, model, n: int, h: int, w: int, device):
        em = Emitter(model, n, device)
        self.em = em
        self.model = model
        spec = model.spec
        dev = device
        self.n = n
        self.x_in = th.empty((n, model.in_channels, h, w), dtype=th.float32, device=dev)
        self.t_in = th.empty((n,), dtype=th.float32, device=dev)
        self.logits = th.empty((n, model.out_channels), dtype=th.float32, device=dev)
        self.dlogits = th.empty((n, model.out_channels), dtype=th.float32, device=dev)
        self.dx = th.empty((n, model.in_channels, h, w), dtype=th.float32, device=dev)
        self.grad_scale = th.ones((), dtype=th.float32)
        film_all = em.emit_embedding(self.t_in, None, spec)
        tape: list = []
        cur: Optional[View] = None
        hh, ww = h, w
        for blk in spec.input_blocks + [spec.middle_block]:
            for l in blk:
                if isinstance(l, ConvInSpec):
                    o = em.act(n, hh, ww, l.cout)
                    em.conv_in(l, self.x_in, o)
                    tape.append(("conv_in", l, o))
                elif isinstance(l, ResSpec):
                    if l.mode == "down":
                        hh, ww = hh // 2, ww // 2
                    o = em.act(n, hh, ww, l.cout)
                    em.res_block(l, cur, o, film_all, tape)
                elif isinstance(l, ResampleSpec):  # classifier_resblock_updown=False: Downsample.op, 3x3 stride 2
                    assert l.mode == "down"
                    hh, ww = (hh - 1) // 2 + 1, (ww - 1) // 2 + 1
                    o = em.act(n, hh, ww, l.ch)
                    em.resample(l, cur, o)
                    tape.append(("down", l, cur, o))
                else:
                    o = em.act(n, hh, ww, l.ch)
                    em.attn_block(l, cur, o, tape)
                cur = o
        # head: GN -> SiLU -> AttentionPool2d (unet.py:833-841)
        st = em.stats_buf()
        pooled_in = em.act(n, hh, ww, cur.c)
        em.gn_stats(cur, st)
        em.gn_apply(cur, st, em.f32("out.0.weight"), em.f32("out.0.bias"), pooled_in, silu=True)
        hw, cch, heads = hh * ww, cur.c, model.pool_heads
        lib = L.load()
        ws = th.empty(int(lib.gd_attnpool_ws_floats(n, hw + 1, cch)), dtype=th.float32, device=dev)
        pos = em.f32("out.2.positional_embedding")
        wqkv = em.P["out.2.qkv_proj.weight"].float().reshape(3 * cch, cch).contiguous()
        bqkv = em.f32("out.2.qkv_proj.bias")
        wc = em.P["out.2.c_proj.weight"].float().reshape(model.out_channels, cch).contiguous()
        bc = em.f32("out.2.c_proj.bias")
        em.keep += [ws, pos, wqkv, bqkv, wc, bc]
        em.prog.add("gd_attnpool_fwd", C.c_void_p(pooled_in.ptr), pooled_in.ld, _p(pos), _p(wqkv), _p(bqkv), _p(wc),
                    _p(bc), _p(self.logits), _p(ws), n, hw, cch, heads, model.out_channels)
        self.fwd = em.prog

        # ------------------------------------------------------------------ backward program
        em.prog = Program(em.device)
        bw = em
        wqkv_t = wqkv.t().contiguous()
        wc_t = wc.t().contiguous()
        em.keep += [wqkv_t, wc_t]
        d_pool = bw.scratch("g_pool", n, hh, ww, cch)
        bw.prog.add("gd_attnpool_bwd", _p(self.dlogits), _p(wqkv_t), _p(wc_t), _p(ws), C.c_void_p(d_pool.ptr), d_pool.ld,
                    n, hw, cch, heads, model.out_channels, C.c_float(self.LOSS_SCALE))
        g = bw.scratch("gA", n, hh, ww, cch)
        bw.gn_bwd(cur, st, em.f32("out.0.weight"), em.f32("out.0.bias"), d_pool, g, silu=True)
        flip = 0
        for entry in reversed(tape):
            kind = entry[0]
            if kind == "res":
                _, r, x, st1, h1, st2, film_ptr, film_ld, o = entry
                k = r.key
                t1 = bw.scratch("t1", n, o.h, o.w, r.cout)
                bw.conv(g, pack_conv3x3_bwd(em.P[f"{k}.out_layers.3.weight"]), None, r.cout, t1)
                t2 = bw.scratch("t2", n, o.h, o.w, r.cout)
                bw.gn_bwd(h1, st2, em.f32(f"{k}.out_layers.0.weight"), em.f32(f"{k}.out_layers.0.bias"), t1, t2,
                          silu=True, film=film_ptr, film_ld=film_ld)
                t3 = bw.scratch("t3", n, o.h, o.w, r.cin)
                bw.conv(t2, pack_conv3x3_bwd(em.P[f"{k}.in_layers.2.weight"]), None, r.cin, t3)
                if r.has_skip_conv:
                    t4 = bw.scratch("t4", n, o.h, o.w, r.cin)
                    bw.conv(g, pack_1x1_bwd(em.P[f"{k}.skip_connection.weight"]), None, r.cin, t4, taps=1)
                    add, add_mode = t4, L.GN_SAME
                else:
                    add, add_mode = g, (L.GN_AVGPOOL2 if r.mode == "down" else L.GN_SAME)
                if r.mode == "up":
                    raise NotImplementedError("up ResBlock backward is not needed by the encoder classifier")
                flip ^= 1
                g_in = bw.scratch("gB" if flip else "gA", n, x.h, x.w, r.cin)
                bw.gn_bwd(x, st1, em.f32(f"{k}.in_layers.0.weight"), em.f32(f"{k}.in_layers.0.bias"), t3, g_in,
                          silu=True, mode=(L.GN_AVGPOOL2 if r.mode == "down" else L.GN_SAME), add=add,
                          add_mode=add_mode)
                g = g_in
            elif kind == "attn":
                _, a, x, st_a, qkv, att, lse, o = entry
                k = a.key
                d_att = bw.scratch("t1", n, x.h, x.w, a.ch)
                bw.conv(g, pack_1x1_bwd(em.P[f"{k}.proj_out.weight"]), None, a.ch, d_att, taps=1)
                dqkv = bw.scratch("t5", n, x.h, x.w, 3 * a.ch)
                delta = th.empty((n, a.heads, x.h * x.w), dtype=th.float32, device=dev)
                em.keep.append(delta)
                order = L.QKV_NEW if a.new_order else L.QKV_LEGACY
                bw.prog.add("gd_attention_bwd", C.c_void_p(qkv.ptr), qkv.ld, C.c_void_p(att.ptr), att.ld,
                            C.c_void_p(d_att.ptr), d_att.ld, _p(lse), _p(delta), C.c_void_p(dqkv.ptr), dqkv.ld, n,
                            x.h * x.w, a.heads, order)
                d_a = bw.scratch("t2", n, x.h, x.w, a.ch)
                bw.conv(dqkv, pack_1x1_bwd(em.P[f"{k}.qkv.weight"]), None, a.ch, d_a, taps=1)
                flip ^= 1
                g_in = bw.scratch("gB" if flip else "gA", n, x.h, x.w, a.ch)
                bw.gn_bwd(x, st_a, em.f32(f"{k}.norm.weight"), em.f32(f"{k}.norm.bias"), d_a, g_in, silu=False, add=g,
                          add_mode=L.GN_SAME)
                g = g_in
            elif kind == "down":  # dX of the strided conv: dcols = dY x W (GEMM), then the transpose of the gather
                _, l, x, o = entry
                wt = pack_conv3x3(em.P[f"{l.key}.weight"])[:l.ch].t().contiguous()  # [9*C_in][C_out]
                dcols = bw.scratch("dcols", n, o.h, o.w, 9 * l.ch)
                bw.conv(g, wt, None, 9 * l.ch, dcols, taps=1)
                flip ^= 1
                g_in = bw.scratch("gB" if flip else "gA", n, x.h, x.w, l.ch)
                bw.prog.add("gd_col2im3x3_s2_nhwc", C.c_void_p(dcols.ptr), dcols.ld, C.c_void_p(g_in.ptr), g_in.ld, n, x.h,
                            x.w, l.ch)
                g = g_in
            else:  # conv_in: dX in fp32 NCHW; its fp32 epilogue undoes the loss scale and applies the user's scale
                _, l, o = entry
                wk = em.P[f"{l.key}.weight"]
                if l.cin <= NARROW_COUT:  # dX = conv3x3 of dY with the flipped, transposed weights
                    bw.conv3x3_narrow(g, wk.flip(2, 3).permute(1, 0, 2, 3), None, self.dx,
                                      out_scale=1.0 / self.LOSS_SCALE, geom=(n, o.h, o.w))
                else:
                    bw.conv(g, pack_conv3x3_bwd(wk), None, l.cin, self.dx, out_mode=L.OUT_NCHW_F32,
                            out_scale=1.0 / self.LOSS_SCALE, geom=(n, o.h, o.w))
        self.bwd = em.prog
        self._scale_slot = em.last_scale_slot  # the launch whose fp32 epilogue carries out_scale
        self._scale = 1.0
        self.fwd_version = 0

    def set_scale(self, scale: float) -> None:
        """The user's classifier_scale is folded into the LAST launch's fp32 epilogue (out_scale = scale / LOSS_SCALE):
        the fp16 gradient chain always carries exactly LOSS_SCALE x the true gradient, whatever the guidance scale, so a
        large scale (ADM configs use up to 10) cannot overflow fp16 on the way."""
        scale = float(scale)
        if scale == self._scale:
            return
        kind, a, pos = self._scale_slot
        val = scale / self.LOSS_SCALE
        if kind == "desc":
            a.out_scale = val
        else:
            fn, args, name = self.bwd.calls[a]
            self.bwd.calls[a] = (fn, args[:pos] + (C.c_float(val),) + args[pos + 1:], name)
        self._scale = scale

    # -- execution --------------------------------------------------------------------------------
    def forward(self, x, timesteps) -> th.Tensor:
        self.x_in.copy_(x)
        self.t_in.copy_(timesteps)
        self.fwd_version += 1
        self.fwd.run()
        return self.logits

    def backward(self, dlogits: th.Tensor) -> th.Tensor:
        """dlogits: fp32 [n, classes] gradient w.r.t. the logits. Returns d/dx in fp32 NCHW (static buffer)."""
        self.dlogits.copy_(dlogits)
        self.set_scale(1.0)
        self.bwd.run()
        return self.dx

    def guidance(self, x, timesteps, y, scale: float) -> th.Tensor:
        """scale * d/dx sum_b log_softmax(classifier(x,t))[b, y_b]   (scripts/classifier_sample.py:54-61)."""
        self.forward(x, timesteps)
        yy = y.to(th.int64).contiguous()
        if yy.device != self.logits.device:
            raise L.GdError(f"labels live on {yy.device} but the classifier runs on {self.logits.device}")
        with th.cuda.device(self.logits.device):  # launch on the plan's device whatever device is current
            stream = C.c_void_p(th.cuda.current_stream().cuda_stream)
            L.check(L.load().gd_logsoftmax_select_bwd(_p(self.logits), _p(yy), _p(self.dlogits), self.n,
                                                       self.logits.shape[1], C.c_float(1.0), stream),
                    "gd_logsoftmax_select_bwd")
        self.set_scale(scale)
        self.bwd.run()
        return self.dx


class _ClassifierFn(th.autograd.Function):
    @staticmethod
    def forward(ctx, x, timesteps, model):
        n, c, h, w = x.shape
        plan = model.plan(n, h, w, x.device)
        ctx.plan = plan
        out = plan.forward(x.detach(), timesteps).clone()
        ctx.version = plan.fwd_version
        return out

    @staticmethod
    def backward(ctx, dlogits):
        if ctx.version != ctx.plan.fwd_version:
            # the plan keeps ONE set of saved activations: a second forward on the same (batch, resolution) plan has
            # overwritten what this backward needs
            raise L.GdError("classifier backward after another forward of the same shape: the saved activations were "
                            "overwritten; call backward (autograd.grad) before the next forward")
        return ctx.plan.backward(dlogits.float().contiguous()).clone(), None, None


def classifier_apply(model, x: th.Tensor, timesteps: th.Tensor) -> th.Tensor:
    _require_cuda(x.device)
    if x.requires_grad and th.is_grad_enabled():
        return _ClassifierFn.apply(x, timesteps, model)
    n, c, h, w = x.shape
    return model.plan(n, h, w, x.device).forward(x, timesteps).clone()
