"""GPU parity: tcgen05 implicit-GEMM conv (gd_conv_igemm) and the direct first-layer conv against
torch.nn.functional.conv2d in fp32 on fp16-rounded operands.  Tolerance: max-abs error / max-abs reference
<= 4e-3 (fp16 output rounding is 2^-11 relative per element; north_star allows 2e-2)."""
import pytest
import torch as th
import torch.nn.functional as F

from guided_diffusion_clip_b200 import _lib as L
from guided_diffusion_clip_b200.engine import pack_1x1, pack_1x1_bwd, pack_conv3x3, pack_conv3x3_bwd, pack_tap_expand
from tests import gpu_helpers as H

pytestmark = pytest.mark.gpu
TOL = 4e-3


@pytest.fixture(autouse=True)
def _fp32_reference():
    th.backends.cudnn.allow_tf32 = False
    th.backends.cuda.matmul.allow_tf32 = False


def _rand(shape, seed, scale=1.0):
    g = th.Generator().manual_seed(seed)
    return (th.randn(shape, generator=g) * scale).cuda()


def _h(x):  # fp16 round trip
    return x.half().float()


CASES = [
    # n, h, w, cin, cout, taps
    (2, 16, 16, 64, 64, 9),
    (1, 32, 32, 128, 256, 9),
    (3, 8, 8, 128, 512, 9),      # 2 images per tile, odd batch -> masked rows, 2 N tiles
    (2, 16, 16, 192, 576, 1),    # 1x1, BN=192
    (1, 24, 24, 64, 128, 9),     # W not a multiple of the patch width
    (2, 64, 64, 64, 64, 9),      # more tiles than one wave of the TMEM double buffer per CTA on small grids
    (1, 16, 16, 512, 128, 9),    # 72 K blocks -> smem ring wraps many times
    (8, 64, 64, 128, 128, 9),    # 256 tiles > 148 SMs: persistent loop + accumulator ping-pong
    (2, 4, 4, 64, 64, 9),        # 4x4 images: 8 images per tile
    (1, 16, 32, 64, 64, 9),      # non-square, halo tiles: 2 patches across
    (2, 32, 16, 128, 128, 9),    # non-square the other way; one patch across, 4 down
    (1, 20, 24, 64, 64, 9),      # neither H nor W a multiple of the 16x8 patch: halo rows and columns run off the image
    (1, 16, 16, 320, 64, 9),     # 5 channel blocks: the 3-slot activation ring wraps out of step with the weight ring
    (2, 16, 8, 64, 128, 9),      # 8-pixel-wide images: halo tile 18 rows x 8 pixels, 1024-byte tap stride
    (5, 32, 32, 64, 64, 9),      # odd tile count: the last CTA pair is half empty
]


@pytest.mark.parametrize("n,h,w,cin,cout,taps", CASES)
def test_conv_matches_torch(lib, n, h, w, cin, cout, taps):
    k = 3 if taps == 9 else 1
    x = _h(_rand((n, cin, h, w), 1))
    wt = _h(_rand((cout, cin, k, k), 2, (cin * k * k) ** -0.5))
    b = _rand((cout,), 3, 0.1)
    ref = F.conv2d(x, wt, b, padding=k // 2)
    pack = pack_conv3x3(wt) if taps == 9 else pack_1x1(wt)
    out = H.conv_igemm(H.nhwc_half(x), cin, 0, pack, b, cout, n, h, w, taps=taps)
    th.cuda.synchronize()
    err = H.rel_err(out.permute(0, 3, 1, 2), ref)
    print(f"conv n={n} {h}x{w} {cin}->{cout} taps={taps}: rel err {err:.3e}")
    assert err < TOL


def test_conv_strided_views_and_fused_skip(lib):
    """K = 9*C0 taps of source 0 + C1 channels of a 1x1 source 1 (ResBlock out conv + skip_connection,
    unet.py:211,222,256), reading/writing channel slices of wider buffers (the free concat)."""
    n, h, w, c0, c1, cout = 2, 16, 16, 64, 128, 128
    a = _h(_rand((n, c0, h, w), 4))
    s = _h(_rand((n, c1, h, w), 5))
    w3 = _h(_rand((cout, c0, 3, 3), 6, (c0 * 9) ** -0.5))
    w1 = _h(_rand((cout, c1, 1, 1), 7, c1 ** -0.5))
    b = _rand((cout,), 8, 0.1)
    ref = F.conv2d(a, w3, b, padding=1) + F.conv2d(s, w1)
    a_buf = H.nhwc_half(a, ld=c0 + 64, off=64)
    s_buf = H.nhwc_half(s, ld=c1 + 192, off=128)
    out = H.conv_igemm(a_buf, c0, 64, pack_conv3x3(w3, w1), b, cout, n, h, w, a1_buf=s_buf, c1=c1, off1=128,
                       ld_out=cout + 64, out_off=64)
    th.cuda.synchronize()
    err = H.rel_err(out[..., 64:].permute(0, 3, 1, 2), ref)
    print(f"fused skip + strided: rel err {err:.3e}")
    assert err < TOL
    assert float(out[..., :64].abs().max()) == 0.0  # nothing written outside the slice


@pytest.mark.parametrize("mode", [L.RES_SAME, L.RES_UPSAMPLE2, L.RES_AVGPOOL2])
def test_conv_residual_modes(lib, mode):
    n, h, w, c = 2, 16, 16, 64
    x = _h(_rand((n, c, h, w), 9))
    wt = _h(_rand((c, c, 3, 3), 10, (c * 9) ** -0.5))
    b = _rand((c,), 11, 0.1)
    if mode == L.RES_SAME:
        r = _h(_rand((n, c, h, w), 12))
        rr = r
    elif mode == L.RES_UPSAMPLE2:
        r = _h(_rand((n, c, h // 2, w // 2), 12))
        rr = F.interpolate(r, scale_factor=2, mode="nearest")
    else:
        r = _h(_rand((n, c, h * 2, w * 2), 12))
        rr = F.avg_pool2d(r, 2)
    ref = F.conv2d(x, wt, b, padding=1) + rr
    out = H.conv_igemm(H.nhwc_half(x), c, 0, pack_conv3x3(wt), b, c, n, h, w, res_buf=H.nhwc_half(r), res_mode=mode)
    th.cuda.synchronize()
    err = H.rel_err(out.permute(0, 3, 1, 2), ref)
    print(f"residual mode {mode}: rel err {err:.3e}")
    assert err < TOL


@pytest.mark.parametrize("cout", [6, 3])
def test_conv_nchw_fp32_small_cout(lib, cout):
    """out head (256->6, unet.py:616) and the final dX conv (128->3): N padded to 16, masked, fp32 NCHW stores."""
    n, h, w, cin = 2, 32, 32, 128
    x = _h(_rand((n, cin, h, w), 13))
    wt = _h(_rand((cout, cin, 3, 3), 14, (cin * 9) ** -0.5))
    b = _rand((cout,), 15, 0.1)
    ref = F.conv2d(x, wt, b, padding=1) * 0.5
    out = H.conv_igemm(H.nhwc_half(x), cin, 0, pack_conv3x3(wt), b, cout, n, h, w, out_mode=L.OUT_NCHW_F32,
                       out_scale=0.5)
    th.cuda.synchronize()
    err = H.rel_err(out, ref)
    print(f"nchw fp32 cout={cout}: rel err {err:.3e}")
    assert err < 1e-3


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 32, 32, 128, 6), (3, 20, 24, 64, 3), (1, 8, 8, 192, 7), (2, 16, 8, 64, 1)])
def test_conv_narrow_output_as_tap_expanded_gemm(lib, n, h, w, cin, cout):
    """The 256->6 head / 128->3 dX convs as ONE 1x1 GEMM with 9*cout columns (engine.pack_tap_expand) + gd_tap_gather3x3,
    against F.conv2d and against the plain 3x3 implicit GEMM; ragged image sizes exercise the zero border of the gather."""
    import ctypes as C
    x = _h(_rand((n, cin, h, w), 31))
    wt = _h(_rand((cout, cin, 3, 3), 32, (cin * 9) ** -0.5))
    b = _rand((cout,), 33, 0.1)
    ref = F.conv2d(x, wt, b, padding=1) * 0.25
    xb = H.nhwc_half(x)
    ytap = H.conv_igemm(xb, cin, 0, pack_tap_expand(wt), None, 64, n, h, w, taps=1)  # fp16 NHWC, 64 columns
    out = th.full((n, cout, h, w), float("nan"), device="cuda")
    L.check(lib.gd_tap_gather3x3(H.vp(ytap), 64, H.vp(b), H.vp(out), n, cout, h, w, C.c_float(0.25), H.stream()))
    plain = H.conv_igemm(xb, cin, 0, pack_conv3x3(wt), b, cout, n, h, w, out_mode=L.OUT_NCHW_F32, out_scale=0.25)
    th.cuda.synchronize()
    err, err2 = H.rel_err(out, ref), H.rel_err(out, plain)
    print(f"narrow conv cout={cout} {h}x{w}: rel err vs torch {err:.3e}, vs plain 3x3 kernel {err2:.3e}")
    assert err < 2e-3 and err2 < 2e-3  # the 9 tap partial sums are rounded to fp16 before the fp32 gather


def test_conv_backward_data_packing(lib):
    """conv backward-data == forward conv with flipped / transposed packed weights (engine.pack_conv3x3_bwd)."""
    n, h, w, cin, cout = 2, 16, 16, 64, 128
    x = _h(_rand((n, cin, h, w), 16)).requires_grad_(True)
    wt = _h(_rand((cout, cin, 3, 3), 17, (cin * 9) ** -0.5))
    dy = _h(_rand((n, cout, h, w), 18))
    F.conv2d(x, wt, padding=1).backward(dy)
    out = H.conv_igemm(H.nhwc_half(dy), cout, 0, pack_conv3x3_bwd(wt), None, cin, n, h, w)
    th.cuda.synchronize()
    err = H.rel_err(out.permute(0, 3, 1, 2), x.grad)
    print(f"bwd-data 3x3: rel err {err:.3e}")
    assert err < TOL
    w1 = _h(_rand((cout, cin, 1, 1), 19, cin ** -0.5))
    x2 = _h(_rand((n, cin, h, w), 20)).requires_grad_(True)
    F.conv2d(x2, w1).backward(dy)
    out = H.conv_igemm(H.nhwc_half(dy), cout, 0, pack_1x1_bwd(w1), None, cin, n, h, w, taps=1)
    th.cuda.synchronize()
    err = H.rel_err(out.permute(0, 3, 1, 2), x2.grad)
    print(f"bwd-data 1x1: rel err {err:.3e}")
    assert err < TOL


def test_conv_linearity_full_size(lib):
    """BASELINE-size property test (256x256, 256->256, batch 2): conv(a*x1 + x2) == a*conv(x1) + conv(x2) up to fp16
    rounding, and a spot check of 4096 random outputs against fp64 dot products."""
    n, h, w, c = 2, 256, 256, 256
    g = th.Generator(device="cuda").manual_seed(21)
    x1 = th.randn((n, h, w, c), generator=g, device="cuda", dtype=th.float16)
    wt = _h(_rand((c, c, 3, 3), 22, (c * 9) ** -0.5))
    pack = pack_conv3x3(wt)
    o1 = H.conv_igemm(x1, c, 0, pack, None, c, n, h, w).float()
    o2 = H.conv_igemm(x1 * 2, c, 0, pack, None, c, n, h, w).float()
    th.cuda.synchronize()
    assert H.rel_err(o2, 2 * o1) < 2e-3
    gi = th.Generator().manual_seed(23)
    idx = th.stack([th.randint(0, n, (4096,), generator=gi), th.randint(0, h, (4096,), generator=gi),
                    th.randint(0, w, (4096,), generator=gi), th.randint(0, c, (4096,), generator=gi)], 1)
    xp = F.pad(x1.double(), (0, 0, 1, 1, 1, 1))
    wd = wt.double()
    worst = 0.0
    for b, y, x, co in idx.tolist()[:512]:
        patch = xp[b, y:y + 3, x:x + 3, :]  # [3,3,C]
        ref = float((patch * wd[co].permute(1, 2, 0)).sum())
        worst = max(worst, abs(ref - float(o1[b, y, x, co])))
    print(f"full-size spot check: worst abs err {worst:.3e} (outputs ~N(0,1))")
    assert worst < 2e-2


@pytest.mark.parametrize("cin", [3, 6])
def test_conv_first_layer_im2col(lib, cin):
    """input_blocks.0.0 (unet.py:483): fp32 NCHW -> 64-wide im2col -> K=64 GEMM; cin=6 is the super-resolution input."""
    from guided_diffusion_clip_b200.engine import pack_conv_in
    n, h, w, cout = 2, 32, 32, 128
    x = _rand((n, cin, h, w), 24)
    wt = _h(_rand((cout, cin, 3, 3), 25, (cin * 9) ** -0.5))
    b = _rand((cout,), 26, 0.1)
    ref = F.conv2d(_h(x), wt, b, padding=1)
    cols = th.zeros((n, h, w, 64), dtype=th.float16, device="cuda")
    L.check(lib.gd_im2col3x3_small_cin(H.vp(x), H.vp(cols), 64, n, cin, h, w, H.stream()))
    out = H.conv_igemm(cols, 64, 0, pack_conv_in(wt), b, cout, n, h, w, taps=1)
    th.cuda.synchronize()
    err = H.rel_err(out.permute(0, 3, 1, 2), ref)
    print(f"first conv cin={cin}: rel err {err:.3e}")
    assert err < TOL


@pytest.mark.parametrize("n,h,w,c_a,c_b", [(2, 32, 32, 128, 128), (3, 8, 8, 256, 128), (2, 16, 16, 128, 384)])
def test_conv_fused_groupnorm_statistics(lib, n, h, w, c_a, c_b):
    """The conv epilogue's per-row-block partial sums + gd_groupnorm_finalize_partials == GroupNorm32 statistics of the
    stored tensor, for one producer and for a skip concatenation of two producers whose groups straddle the boundary."""
    import ctypes as C
    cin = 64
    x = _h(_rand((n, cin, h, w), 30))
    rpi = C.c_int32(0)
    rows = int(lib.gd_conv_stats_rows(n, h, w, C.byref(rpi)))
    assert rpi.value > 0 and rows >= rpi.value * n
    cat = th.zeros((n, h, w, c_a + c_b), dtype=th.float16, device="cuda")
    parts = []
    for off, c, seed in ((0, c_a, 31), (c_a, c_b, 32)):
        wt = _h(_rand((c, cin, 3, 3), seed, (cin * 9) ** -0.5))
        b = _rand((c,), seed + 10, 0.5)
        st = th.zeros((rows, c // 4, 2), dtype=th.float32, device="cuda")
        H.conv_igemm(H.nhwc_half(x), cin, 0, pack_conv3x3(wt), b, c, n, h, w, ld_out=c_a + c_b, out_off=off, out_buf=cat,
                     stats_out=st)
        parts.append(st)
    th.cuda.synchronize()
    stored = cat.float().permute(0, 3, 1, 2)  # what a later GroupNorm will actually read

    def ref_stats(t):
        g = t.reshape(n, 32, -1)
        return g.mean(-1), (g.var(-1, unbiased=False) + 1e-5).rsqrt()

    for (p0, c0, p1, c1, t) in ((parts[0], c_a, None, 0, stored[:, :c_a]), (parts[0], c_a, parts[1], c_b, stored)):
        out = th.zeros((n, 32, 2), dtype=th.float32, device="cuda")
        L.check(lib.gd_groupnorm_finalize_partials(H.vp(p0), c0, c0 // 4, H.vp(p1), c1, c1 // 4, rpi.value, n, h * w,
                                                   C.c_float(1e-5), H.vp(out), None, None, None, 0, None, H.stream()))
        th.cuda.synchronize()
        mean, rstd = ref_stats(t)
        assert float((out[..., 0] - mean).abs().max()) < 2e-4
        assert H.rel_err(out[..., 1], rstd) < 2e-4


def test_conv_rejects_bad_arguments(lib):
    d = L.ConvDesc()
    rc = lib.gd_conv_igemm(d, None)
    assert rc != 0 and b"null" in lib.gd_last_error()


@pytest.mark.parametrize("n,h,w,cin,cout,ld,off", [(2, 16, 16, 3, 64, 64, 0), (3, 32, 32, 3, 256, 512, 256),
                                                   (2, 64, 64, 6, 192, 192, 0), (1, 16, 8, 3, 128, 136, 8),
                                                   (2, 24, 16, 4, 64, 64, 0)])
def test_conv_in3x3_direct_first_layer(lib, n, h, w, cin, cout, ld, off):
    """gd_conv_in3x3 (first layer in one launch from the fp32 NCHW input) against F.conv2d, written into a channel
    slice of a wider buffer; its fused GroupNorm partials, finalized by gd_groupnorm_finalize_partials, against the
    statistics kernel on the stored tensor."""
    import ctypes as C
    from guided_diffusion_clip_b200.engine import pack_conv_in
    x = _rand((n, cin, h, w), 41)
    wt = _h(_rand((cout, cin, 3, 3), 42, (cin * 9) ** -0.5))
    b = _rand((cout,), 43, 0.1)
    ref = F.conv2d(_h(x), wt, b, padding=1)
    wp = pack_conv_in(wt)
    out = th.zeros((n, h, w, ld), dtype=th.float16, device="cuda")
    rpi = C.c_int32(0)
    rows = int(lib.gd_conv_stats_rows(n, h, w, C.byref(rpi)))
    assert rows > 0
    part = th.zeros((rows, cout // 4, 2), device="cuda")
    d = L.ConvInDesc()
    d.x, d.wpack, d.bias, d.out, d.stats_out = x.data_ptr(), wp.data_ptr(), b.data_ptr(), out.data_ptr() + 2 * off, part.data_ptr()
    d.n, d.cin, d.h, d.w, d.cout, d.ld_out = n, cin, h, w, cout, ld
    L.check(lib.gd_conv_in3x3(C.byref(d), H.stream()), "gd_conv_in3x3")
    th.cuda.synchronize()
    got = out[..., off:off + cout].permute(0, 3, 1, 2).float()
    err = H.rel_err(got, ref)
    print(f"conv_in3x3 cin={cin} cout={cout} {h}x{w}: rel err {err:.3e}")
    assert err < 2e-3
    if off:
        assert float(out[..., :off].abs().max()) == 0.0
    if off + cout < ld:
        assert float(out[..., off + cout:].abs().max()) == 0.0
    if (cout // 32) % 4:
        return  # the finalize kernel sums whole 4-channel chunks per group: groups of 2 channels use gd_groupnorm_stats
    st = th.empty((n, 32, 2), device="cuda")
    g = th.Generator().manual_seed(5)
    gamma, beta = (1 + 0.1 * th.randn(cout, generator=g)).cuda(), (0.1 * th.randn(cout, generator=g)).cuda()
    film = (0.2 * th.randn(n, 2 * cout + 4, generator=g)).cuda()
    coef = th.empty((n, cout // 8, 16), device="cuda")
    L.check(lib.gd_groupnorm_finalize_partials(H.vp(part), cout, cout // 4, None, 0, 0, rpi.value, n, h * w,
                                               C.c_float(1e-5), H.vp(st), H.vp(gamma), H.vp(beta), H.vp(film),
                                               film.shape[1], H.vp(coef), H.stream()))
    st_ref = H.gn_stats(out, cout, off)
    th.cuda.synchronize()
    assert float((st[..., 0] - st_ref[..., 0]).abs().max()) < 1e-4
    assert H.rel_err(st[..., 1], st_ref[..., 1]) < 1e-4
    # the affine table written next to the statistics == the standalone table kernel on the same statistics, bit for bit
    assert th.equal(coef, H.gn_coef(st, gamma, beta, film, n, cout))


# ---- GroupNorm (+FiLM, +SiLU, + nearest x2) fused into the conv's operand path -----------------------------------
GN_FUSED_CASES = [
    # n, h, w, c0, cout, c1 (fused 1x1 skip), film, residual, upsample
    (3, 32, 32, 128, 128, 0, True, False, False),    # N = 128 tiles (6-slot ring), odd tile count in pair mode
    (2, 64, 64, 256, 256, 0, True, True, False),     # N = 256 tiles (4-slot ring), identity residual
    (2, 24, 40, 64, 128, 0, False, False, False),    # neither H nor W a multiple of the 16 x 8 tile
    (2, 32, 32, 192, 64, 128, True, False, False),   # fused 1x1-skip operand shares the activation ring (TMA slots)
    (2, 32, 32, 128, 128, 0, True, True, True),      # "up" block: source 16 x 16, nearest x2 folded into the gather
    (1, 8, 16, 64, 64, 0, False, False, False),      # a single tile: one CTA, no pair
    (1, 16, 16, 320, 64, 0, True, False, False),     # 5 channel blocks: ring positions wrap out of step
    (1, 48, 16, 64, 576, 64, True, False, False),    # 3 N tiles of 192 columns per pixel tile + skip operand
    (4, 128, 128, 256, 256, 0, True, True, False),   # 512 tiles: every CTA pair walks several tiles (4-slot ring)
    (2, 128, 128, 128, 256, 192, True, False, False),  # several tiles per pair AND a 3-block skip operand in the ring
    (3, 64, 64, 64, 128, 320, False, False, False),  # skip operand longer than the ring (5 blocks, 6 slots)
]


@pytest.mark.parametrize("n,h,w,c0,cout,c1,film,res,up", GN_FUSED_CASES)
def test_conv_fused_groupnorm_operand_is_bit_identical_to_apply_then_conv(lib, n, h, w, c0, cout, c1, film, res, up):
    """conv3x3(pad0(SiLU(FiLM(GN(x))))) with the normalisation done inside the conv's operand path must equal
    gd_groupnorm_apply followed by the plain conv BIT FOR BIT (same affine, same activation, same fp16 rounding, same
    MMA order), including the zero padding of the NORMALISED tensor at the image border (unet.py:184-185, 205-211)."""
    assert lib.gd_conv_gn_fusable(h, w) == 1
    hs, ws = (h // 2, w // 2) if up else (h, w)
    x = _h(_rand((n, c0, hs, ws), 21, 1.5)) + 0.3
    x_buf = H.nhwc_half(x, ld=c0 + 64, off=32)          # a channel slice of a wider buffer
    gamma = (1.0 + 0.2 * _rand((c0,), 22)).contiguous()
    beta = (0.2 * _rand((c0,), 23)).contiguous()
    fl = (0.3 * _rand((n, 2 * c0 + 8), 24)).contiguous() if film else None
    x_view = x_buf[..., 32:32 + c0]
    st = H.gn_stats(x_buf, c0, off=32)
    coef = H.gn_coef(st, gamma, beta, fl, n, c0)
    wt = _h(_rand((cout, c0, 3, 3), 25, (c0 * 9) ** -0.5))
    w1 = _h(_rand((cout, c1, 1, 1), 26, c1 ** -0.5)) if c1 else None
    s_buf = H.nhwc_half(_h(_rand((n, c1, h, w), 27)), ld=c1 + 64, off=64) if c1 else None
    r_buf = H.nhwc_half(_h(_rand((n, cout, hs, ws), 28))) if res else None
    b = _rand((cout,), 29, 0.1)
    pack = pack_conv3x3(wt, w1)
    kw = dict(a1_buf=s_buf, c1=c1, off1=64 if c1 else 0, res_buf=r_buf,
              res_mode=(L.RES_UPSAMPLE2 if up else L.RES_SAME) if res else L.RES_NONE)
    normed = H.gn_apply(x_buf, c0, st, gamma, beta, film=fl, silu=True, mode=L.GN_UPSAMPLE2 if up else L.GN_SAME, off=32)
    two = H.conv_igemm(normed, c0, 0, pack, b, cout, n, h, w, **kw)
    fused = H.conv_igemm(x_buf, c0, 32, pack, b, cout, n, h, w,
                         gn=dict(mode=L.CONV_GN_UPSAMPLE2 if up else L.CONV_GN_SAME, silu=True, coef=coef), **kw)
    th.cuda.synchronize()
    assert th.isfinite(fused.float()).all()
    diff = (fused.float() - two.float()).abs().max().item()
    print(f"fused GN conv n={n} {h}x{w} {c0}->{cout} skip={c1} film={film} res={res} up={up}: max |diff| {diff}")
    assert th.equal(fused, two)
    # and against torch in fp32 (the two-kernel path is already covered; this guards the test itself)
    xn = F.group_norm(x, 32, gamma, beta, eps=1e-5)
    if film:
        xn = xn * (1 + fl[:, :c0, None, None]) + fl[:, c0:2 * c0, None, None]
    xn = F.silu(xn)
    if up:
        xn = F.interpolate(xn, scale_factor=2, mode="nearest")
    ref = F.conv2d(_h(xn), wt, b, padding=1)
    assert H.rel_err(fused.permute(0, 3, 1, 2), ref) < 2e-2 or c1 or res


def test_conv_fused_groupnorm_rejects_ineligible_geometry(lib):
    assert lib.gd_conv_gn_fusable(8, 8) == 0 and lib.gd_conv_gn_fusable(16, 8) == 0 and lib.gd_conv_gn_fusable(64, 64) == 1
    x_buf = th.zeros((2, 8, 8, 64), dtype=th.float16, device="cuda")
    with pytest.raises(L.GdError, match="fusable"):
        H.conv_igemm(x_buf, 64, 0, pack_conv3x3(th.zeros(64, 64, 3, 3)).cuda(), None, 64, 2, 8, 8,
                     gn=dict(mode=L.CONV_GN_SAME, coef=th.zeros(2, 8, 16, device="cuda")))


SPLITK_CASES = [
    # n, h, w, cin, cout, c1, res: few pixel tiles, long K (the 8x8 / 16x16 layers at per-GPU batch 8, unet.py:552-609)
    (8, 8, 8, 512, 512, 0, False),     # classifier 8x8: 2 pair tiles x 4 N tiles -> 8 splits of one channel block
    (8, 8, 8, 1024, 1024, 0, True),    # UNet 8x8 with residual
    (3, 8, 8, 320, 128, 0, True),      # odd batch (half-empty tile, masked rows), 5 channel blocks -> uneven splits
    (1, 8, 8, 256, 256, 0, False),     # one pixel tile: single-CTA kernel
    (4, 16, 16, 512, 256, 0, True),    # one image per tile, two tiles per image: one partial row per tile
    (2, 8, 8, 256, 256, 128, False),   # fused 1x1-skip operand split along with the main channel blocks
    (5, 16, 8, 128, 128, 0, False),    # 8-pixel-wide halo tiles
]


@pytest.mark.parametrize("n,h,w,cin,cout,c1,res", SPLITK_CASES)
def test_conv_split_k_matches_unsplit_and_torch(lib, n, h, w, cin, cout, c1, res):
    """gd_conv_desc.splitk_ws: K split over idle SMs + fixed-order fp32 reduction.  Output within fp16 rounding of the
    unsplit launch and of torch; fused GroupNorm partials give the same statistics; replays are bit-identical."""
    if not hasattr(lib, "gd_debug_set"):
        pytest.skip("library built without GD_B200_DEVTOOLS")
    lib.gd_debug_set(9, 1)  # split-K is opt-in (GD_B200_SPLITK=1): it trades batch invariance of the low-order bits
    try:
        _split_k_case(lib, n, h, w, cin, cout, c1, res)
    finally:
        lib.gd_debug_set(9, 0)


def _split_k_case(lib, n, h, w, cin, cout, c1, res):
    import ctypes as C
    x = _h(_rand((n, cin, h, w), 50))
    wt = _h(_rand((cout, cin, 3, 3), 51, (cin * 9) ** -0.5))
    b = _rand((cout,), 52, 0.3)
    ref = F.conv2d(x, wt, b, padding=1)
    kw = {}
    w1 = None
    if c1:
        s = _h(_rand((n, c1, h, w), 53))
        w1 = _h(_rand((cout, c1, 1, 1), 54, c1 ** -0.5))
        ref = ref + F.conv2d(s, w1)
        kw.update(a1_buf=H.nhwc_half(s), c1=c1)
    if res:
        r = _h(_rand((n, cout, h, w), 55))
        ref = ref + r
        kw.update(res_buf=H.nhwc_half(r), res_mode=L.RES_SAME)
    pack = pack_conv3x3(wt, w1)
    rpi = C.c_int32(0)
    rows = int(lib.gd_conv_stats_rows(n, h, w, C.byref(rpi)))
    xb = H.nhwc_half(x)

    d = L.ConvDesc()
    d.a0, d.c0, d.ld0, d.taps, d.n, d.h, d.w = xb.data_ptr(), cin, cin, 9, n, h, w
    d.k_total, d.n_pad, d.cout, d.out_mode, d.ld_out = pack.shape[1], cout, cout, L.OUT_NHWC_F16, cout
    d.out = xb.data_ptr()  # only alignment is inspected
    d.bias = b.data_ptr()
    need = int(lib.gd_conv_splitk_ws_bytes(C.byref(d)))
    assert need > 0, "this geometry is expected to split"
    outs, stats = [], []
    for ws_bytes in (0, need, need, need // 2):
        st = th.zeros((rows, cout // 4, 2), dtype=th.float32, device="cuda")
        ws = th.full((ws_bytes // 4,), float("nan"), dtype=th.float32, device="cuda") if ws_bytes else None
        lib.gd_launch_count_reset()
        o = H.conv_igemm(xb, cin, 0, pack, b, cout, n, h, w, stats_out=st, splitk_ws=ws, **kw)
        th.cuda.synchronize()
        launches = int(lib.gd_launch_count())
        assert launches == (2 if ws_bytes == need else launches)  # the full workspace must split
        outs.append(o)
        stats.append(st)
    assert int(lib.gd_launch_count()) >= 1
    err = H.rel_err(outs[1].permute(0, 3, 1, 2), ref)
    err01 = H.rel_err(outs[1].float(), outs[0].float())
    print(f"split-K n={n} {h}x{w} {cin}(+{c1})->{cout} res={res}: vs torch {err:.3e}, vs unsplit {err01:.3e}")
    assert err < TOL and err01 < 2e-3
    assert H.rel_err(outs[3].permute(0, 3, 1, 2), ref) < TOL   # a smaller workspace: fewer splits (or none), same result
    assert th.equal(outs[1], outs[2]) and th.equal(stats[1], stats[2])  # fixed summation order
    for st in (stats[1], stats[3]):
        out_a = th.zeros((n, 32, 2), dtype=th.float32, device="cuda")
        out_b = th.zeros((n, 32, 2), dtype=th.float32, device="cuda")
        for src, dst in ((stats[0], out_a), (st, out_b)):
            L.check(lib.gd_groupnorm_finalize_partials(H.vp(src), cout, cout // 4, None, 0, 0, rpi.value, n, h * w,
                                                       C.c_float(1e-5), H.vp(dst), None, None, None, 0, None, H.stream()))
        th.cuda.synchronize()
        g = outs[1].float().permute(0, 3, 1, 2).reshape(n, 32, -1)
        assert float((out_b[..., 0] - g.mean(-1)).abs().max()) < 3e-4
        assert H.rel_err(out_b[..., 1], (g.var(-1, unbiased=False) + 1e-5).rsqrt()) < 3e-4
        assert H.rel_err(out_b, out_a) < 3e-4


def test_conv_split_k_is_not_used_for_large_grids(lib):
    """Enough pixel tiles to fill the GPU: the workspace is ignored (one launch)."""
    import ctypes as C
    n, h, w, c = 8, 64, 64, 128
    x = _h(_rand((n, c, h, w), 60))
    wt = _h(_rand((c, c, 3, 3), 61, (c * 9) ** -0.5))
    xb = H.nhwc_half(x)
    d = L.ConvDesc()
    d.a0, d.c0, d.ld0, d.taps, d.n, d.h, d.w = xb.data_ptr(), c, c, 9, n, h, w
    d.k_total, d.n_pad, d.cout, d.out_mode, d.ld_out, d.out = 9 * c, c, c, L.OUT_NHWC_F16, c, xb.data_ptr()
    if hasattr(lib, "gd_debug_set"):
        lib.gd_debug_set(9, 1)
    try:
        assert int(lib.gd_conv_splitk_ws_bytes(C.byref(d))) == 0
        ws = th.empty(1 << 20, dtype=th.float32, device="cuda")
        lib.gd_launch_count_reset()
        out = H.conv_igemm(xb, c, 0, pack_conv3x3(wt), None, c, n, h, w, splitk_ws=ws)
        th.cuda.synchronize()
        assert int(lib.gd_launch_count()) == 1
    finally:
        if hasattr(lib, "gd_debug_set"):
            lib.gd_debug_set(9, 0)
    assert H.rel_err(out.permute(0, 3, 1, 2), F.conv2d(x, wt, None, padding=1)) < TOL
