"""CPU: the product's host-side logic against the reference golden vectors — respaced timestep sets / maps
(bit-exact integers), float64 tables, factory defaults, state_dict layout — plus the C-ABI surface:
libgd_b200.so loads, exports every symbol include/gd_b200.h declares, validates arguments without a GPU and
the package refuses to compute on CPU tensors (no fallback)."""
import ctypes
import hashlib
import json
import os
import re

import numpy as np
import pytest
import torch as th

from guided_diffusion_clip_b200 import gaussian_diffusion as gd
from guided_diffusion_clip_b200 import respace
from guided_diffusion_clip_b200 import script_util as su
from oracle import golden_cfg as cfg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def J(golden_dir):
    with open(os.path.join(golden_dir, "diffusion_golden.json")) as f:
        return json.load(f)


def test_space_timesteps_bit_exact(J):
    for key, want in J["space_timesteps"].items():
        T, spec = key.split("|")
        got = respace.space_timesteps(int(T), spec)
        assert isinstance(got, set) and sorted(got) == want, key
    assert sorted(respace.space_timesteps(300, [10, 15, 20])) == J["space_timesteps"]["300|10,15,20"]
    for key, msg in J["errors"].items():
        T, spec = key.split("|")
        with pytest.raises(ValueError) as e:
            respace.space_timesteps(int(T), spec)
        assert str(e.value) == msg


def test_survey_appendix_c_vectors():
    def h(xs):
        return hashlib.sha256(np.array(sorted(xs), dtype="<i8").tobytes()).hexdigest()[:16]
    assert h(respace.space_timesteps(1000, "25")) == "e22cac1bf1562a06"
    assert h(respace.space_timesteps(1000, "250")) == "d802631b9565348c"
    assert h(respace.space_timesteps(1000, "ddim25")) == "5ac9828e4fbaa105"
    assert h(respace.space_timesteps(1000, "ddim50")) == "ebc60f0baaa7a6bc"
    assert h(respace.space_timesteps(1000, "50")) == "c5bdf9b959c7973b"
    assert h(respace.space_timesteps(1000, "100")) == "432f07a847b37dff"
    assert h(respace.space_timesteps(300, "10,15,20")) == "9293baff71d3533b"


def test_diffusion_tables_and_maps_bit_exact(J):
    for name, kw in cfg.DIFFUSION_CASES.items():
        d = su.create_gaussian_diffusion(**kw)
        g = J["tables"][name]
        assert d.timestep_map == J["maps"][name]
        assert d.num_timesteps == g["num_timesteps"]
        assert d.model_mean_type.name == g["model_mean_type"] and d.model_var_type.name == g["model_var_type"]
        assert d.loss_type.name == g["loss_type"] and bool(d.rescale_timesteps) == g["rescale_timesteps"]
        for key in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod",
                    "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
                    "posterior_mean_coef1", "posterior_mean_coef2"):
            a, b = getattr(d, key), np.array(g[key], dtype=np.float64)
            assert a.dtype == np.float64 and np.array_equal(a, b), (name, key)  # same float64 bits
        tab = d.coef_table()
        assert tab.dtype == np.float32 and tab.shape == (d.num_timesteps, 12)
        assert tab[0, 10] == 0.0 and (tab[1:, 10] == 1.0).all()


def test_factory_defaults_match_reference(J):
    D = J["defaults"]
    assert su.diffusion_defaults() == D["diffusion_defaults"]
    assert su.classifier_defaults() == D["classifier_defaults"]
    assert su.model_and_diffusion_defaults() == D["model_and_diffusion_defaults"]
    assert su.sr_model_and_diffusion_defaults() == D["sr_model_and_diffusion_defaults"]
    assert su.NUM_CLASSES == 512
    assert su.str2bool("yes") and not su.str2bool("0")


def test_wrapped_model_maps_and_rescales():
    d = su.create_gaussian_diffusion(steps=1000, learn_sigma=True, timestep_respacing="10", rescale_timesteps=True)
    seen = {}
    w = d._wrap(lambda x, t, **kw: seen.setdefault("t", t))
    w(None, th.tensor([9, 0, 3]))
    assert seen["t"].dtype == th.float32 and seen["t"].tolist() == [999.0, 0.0, 333.0]
    assert d._wrap(w) is w
    d2 = su.create_gaussian_diffusion(steps=4000, learn_sigma=True, timestep_respacing="ddim25")
    got = {}
    d2._wrap(lambda x, t, **kw: got.setdefault("t", t))(None, th.tensor([24, 1]))
    assert got["t"].dtype == th.int64 and got["t"].tolist() == [3840, 160]


def _layout_hash(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(f"{k}:{tuple(v.shape)};".encode())
    return h.hexdigest()[:16], len(sd)


def test_state_dict_layout_matches_reference(J):
    """Key names, order and shapes of SURVEY App. E; hashes recorded from the reference modules by make_golden.py."""
    L = J["layouts"]
    m = su.create_model(**cfg.UNET_KW)
    assert list(_layout_hash(m.state_dict())) == L["unet_tiny"]
    c = su.create_classifier(**cfg.CLASSIFIER_KW)
    assert list(_layout_hash(c.state_dict())) == L["clf_tiny"]
    with th.device("meta"):
        m256 = su.create_model(**cfg.UNET256_KW)
        c256 = su.create_classifier(**cfg.CLF256_KW)
        sr = su.sr_create_model(**cfg.SR512_KW)
    assert list(_layout_hash(m256.state_dict())) == L["unet_256"] and len(m256.state_dict()) == 567
    assert list(_layout_hash(c256.state_dict())) == L["clf_256"] and len(c256.state_dict()) == 249
    assert list(_layout_hash(sr.state_dict())) == L["sr_512"]
    assert sum(p.numel() for p in m256.parameters()) == L["unet_256_params"]


def test_zero_modules_and_fp16_conversion():
    m = su.create_model(**cfg.UNET_KW)
    sd = m.state_dict()
    assert float(sd["out.2.weight"].abs().max()) == 0.0                       # unet.py:616
    assert float(sd["input_blocks.1.0.out_layers.3.weight"].abs().max()) == 0  # unet.py:210
    assert float(sd["middle_block.1.proj_out.weight"].abs().max()) == 0        # unet.py:294
    m.convert_to_fp16()
    sd = m.state_dict()
    assert sd["middle_block.1.qkv.weight"].dtype == th.float16
    assert sd["middle_block.0.in_layers.0.weight"].dtype == th.float32   # GroupNorm stays fp32
    assert sd["time_embed.0.weight"].dtype == th.float32
    m.convert_to_fp32()
    assert m.state_dict()["middle_block.1.qkv.weight"].dtype == th.float32


def test_unsupported_configurations_raise():
    with pytest.raises(NotImplementedError):
        from guided_diffusion_clip_b200.unet import UNetModel
        UNetModel(64, 3, 64, 6, 1, (4,), conv_resample=False, num_head_channels=64)
    with pytest.raises(ValueError):
        su.create_model(96, 64, 1)
    with pytest.raises(NotImplementedError):
        from guided_diffusion_clip_b200.unet import EncoderUNetModel
        EncoderUNetModel(64, 3, 64, 10, 1, (4,), num_head_channels=64, resblock_updown=True, pool="adaptive")


# ------------------------------------------------------------------------------------------------ C ABI
def _declared_symbols(header="gd_b200.h"):
    with open(os.path.join(ROOT, "include", header)) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    from guided_diffusion_clip_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gd_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names  # the ctypes binding covers exactly the header
    assert lib.gd_version() == 1
    # measurement hooks are a separate header, compiled in only with -DGD_B200_DEVTOOLS, and not in the product ABI
    dev = _declared_symbols("gd_b200_devtools.h")
    assert dev == sorted(_lib.DEV_SIGNATURES) and not set(dev) & set(names)


def test_abi_validates_arguments_without_gpu(lib):
    from guided_diffusion_clip_b200 import _lib
    d = _lib.ConvDesc()
    assert lib.gd_conv_igemm(d, None) == -1
    assert b"null tensor pointer" in lib.gd_last_error()
    d.a0 = d.wpack = d.out = 16
    d.taps, d.c0, d.ld0 = 9, 48, 48
    assert lib.gd_conv_igemm(d, None) == -1 and b"multiple of 64" in lib.gd_last_error()
    assert lib.gd_attention_fwd(ctypes.c_void_p(16), 192, ctypes.c_void_p(16), 64, None, 1, 65, 1, 0, None) == -1
    assert b"multiple of 64" in lib.gd_last_error()
    assert lib.gd_groupnorm_stats(ctypes.c_void_p(16), 48, 1, 4, 48, ctypes.c_float(1e-5), ctypes.c_void_p(16),
                                  ctypes.c_void_p(16), None) == -1
    assert b"multiple of 32" in lib.gd_last_error()
    assert lib.gd_tap_gather3x3(ctypes.c_void_p(16), 64, None, ctypes.c_void_p(16), 1, 8, 4, 4, ctypes.c_float(1.0),
                                None) == -1
    assert b"outside [1,7]" in lib.gd_last_error()
    p = _lib.PosteriorDesc()
    assert lib.gd_posterior_step(p, None) == -1
    assert lib.gd_groupnorm_ws_floats(8, 1, 32) == 8 * 129 * 64


def test_no_cpu_fallback():
    from guided_diffusion_clip_b200._lib import GdError
    m = su.create_model(**cfg.UNET_KW)
    x, t, y = cfg.model_inputs()
    with pytest.raises(GdError):
        m(x, t, y)
    d = su.create_gaussian_diffusion(steps=1000, learn_sigma=True, timestep_respacing="4")
    with pytest.raises(GdError):
        d.p_sample(lambda x, t, **k: th.zeros(2, 6, 8, 8), th.zeros(2, 3, 8, 8), th.tensor([1, 1]), model_kwargs={})
    c = su.create_classifier(**cfg.CLASSIFIER_KW)
    with pytest.raises(GdError):
        c(x, t)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "guided_diffusion_clip_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
                assert "/root/reference" not in src, fn


def test_upsampler_low_res_stream_is_rank_strided(tmp_path):
    """super_res_sample.py:77-100: rank r reads base samples r, r+W, ...; batches wrap around the file with the
    partial batch carried over; low_res = uint8/127.5 - 1 in NCHW; labels follow the same order (bit-exact)."""
    import itertools
    from guided_diffusion_clip_b200 import dist_util
    n = 7
    rng = np.random.RandomState(3)
    arr = rng.randint(0, 256, size=(n, 4, 4, 3)).astype(np.uint8)
    lab = np.arange(100, 100 + n).astype(np.int64)
    path = str(tmp_path / "base.npz")
    np.savez(path, arr, lab)
    for world, r, bs in ((2, 0, 3), (2, 1, 3), (3, 2, 2), (1, 0, 4)):
        mine = list(range(r, n, world))
        order = list(itertools.islice(itertools.cycle(mine), 4 * bs))
        gen = dist_util.load_data_for_worker(path, bs, True, rank_=r, world=world)
        for k in range(4):
            b = next(gen)
            idx = order[k * bs:(k + 1) * bs]
            assert b["y"].dtype == th.int64 and b["y"].tolist() == lab[idx].tolist()
            want = th.from_numpy(arr[idx]).float() / 127.5 - 1.0
            assert b["low_res"].shape == (bs, 3, 4, 4) and th.equal(b["low_res"], want.permute(0, 3, 1, 2))
    gen = dist_util.load_data_for_worker(path, 2, False, rank_=0, world=1)
    assert set(next(gen).keys()) == {"low_res"}
    with pytest.raises(ValueError):
        next(dist_util.load_data_for_worker(path, 2, True, rank_=9, world=10))


def test_weight_packing_algebra_on_cpu():
    """The host-side weight packers restated as plain matrix products (no kernel involved): the tap-expanded packing
    of the narrow-output convs (engine.pack_tap_expand + the gather gd_tap_gather3x3 performs) and the flipped /
    transposed backward-data packing reproduce F.conv2d and its data gradient."""
    import torch.nn.functional as F
    from guided_diffusion_clip_b200.engine import pack_conv3x3, pack_conv3x3_bwd, pack_tap_expand
    g = th.Generator().manual_seed(0)
    n, ci, co, h, w = 2, 8, 3, 6, 5
    x = th.randn(n, ci, h, w, generator=g).half().float()
    wt = th.randn(co, ci, 3, 3, generator=g).half().float()
    ref = F.conv2d(x, wt, padding=1)
    # tap-expanded 1x1 GEMM: y[n,y,x,tap*co+c], then out[p] = sum_tap y[p + delta_tap][tap]
    wp = pack_tap_expand(wt).float()
    assert wp.shape == (64, ci) and float(wp[9 * co:].abs().max()) == 0.0
    ytap = th.einsum("nchw,kc->nhwk", x, wp)
    ypad = F.pad(ytap, (0, 0, 1, 1, 1, 1))
    out = th.zeros(n, co, h, w)
    for ky in range(3):
        for kx in range(3):
            t = ky * 3 + kx
            out += ypad[:, ky:ky + h, kx:kx + w, t * co:(t + 1) * co].permute(0, 3, 1, 2)
    assert th.allclose(out, ref, atol=1e-4)
    # implicit-GEMM packing: k = (ky*3+kx)*Ci + ci over an im2col of the zero-padded input
    cols = F.unfold(x, 3, padding=1).view(n, ci, 9, h * w).permute(0, 3, 2, 1).reshape(n, h * w, 9 * ci)
    out2 = (cols @ pack_conv3x3(wt).float()[:co].t()).permute(0, 2, 1).reshape(n, co, h, w)
    assert th.allclose(out2, ref, atol=1e-4)
    # backward-data = forward conv of dy with the flipped, transposed weights
    xg = x.clone().requires_grad_(True)
    dy = th.randn(n, co, h, w, generator=g).half().float()
    F.conv2d(xg, wt, padding=1).backward(dy)
    dcols = F.unfold(dy, 3, padding=1).view(n, co, 9, h * w).permute(0, 3, 2, 1).reshape(n, h * w, 9 * co)
    dx = (dcols @ pack_conv3x3_bwd(wt).float()[:ci].t()).permute(0, 2, 1).reshape(n, ci, h, w)
    assert th.allclose(dx, xg.grad, atol=1e-4)


def test_fused_statistics_geometry_excludes_interleaved_images(lib):
    """gd_conv_stats_rows: one partial row per 128-pixel tile (images >= 128 pixels per tile), four per tile when a tile
    holds several WHOLE images (8x8), and none when several images share a tile AND an image spans several tiles
    (12x12, 10x10: the rows of one image would interleave with its tile mate's) — found by the 48x48 model test."""
    import ctypes
    rpi = ctypes.c_int32(-1)
    assert lib.gd_conv_stats_rows(4, 256, 256, ctypes.byref(rpi)) == 4 * 512 and rpi.value == 512
    assert lib.gd_conv_stats_rows(4, 8, 8, ctypes.byref(rpi)) == 8 and rpi.value == 2
    assert lib.gd_conv_stats_rows(3, 8, 8, ctypes.byref(rpi)) == 8 and rpi.value == 2   # last pair half empty
    for hw in (12, 10, 6):
        assert lib.gd_conv_stats_rows(4, hw, hw, ctypes.byref(rpi)) == 0 and rpi.value == 0
    assert lib.gd_conv_stats_rows(4, 24, 24, ctypes.byref(rpi)) == 24 and rpi.value == 6


def test_bench_reference_arm_prints_exactly_one_json_line():
    """The driver parses bench.py's stdout: ONE JSON line, whatever libraries print (they are sent to stderr).  The
    reference arm runs on the host cores, so this is testable without a GPU (64x64 keeps it to seconds)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--image-size", "64", "--gpus", "1"], capture_output=True, text=True,
                       timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1


def test_strided_conv_packing_algebra_on_cpu():
    """The host-side algebra behind Downsample.op on the CUDA path, restated on CPU: (a) the stride-2 gather layout
    [.., tap*C + ci] times pack_conv3x3's rows equals F.conv2d(stride=2, padding=1); (b) the data-gradient is the GEMM
    with the transposed pack followed by the transpose of the gather (F.fold), as engine.ClassifierPlan emits it."""
    import torch.nn.functional as F
    from guided_diffusion_clip_b200.engine import pack_conv3x3
    g = th.Generator().manual_seed(3)
    n, c, co, h, w = 2, 16, 32, 9, 12
    x = th.randn(n, c, h, w, generator=g)
    wt = th.randn(co, c, 3, 3, generator=g) / 12
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    cols = F.unfold(x, 3, padding=1, stride=2).reshape(n, c, 9, ho * wo).permute(0, 3, 2, 1).reshape(n, ho * wo, 9 * c)
    wp = pack_conv3x3(wt)[:co].float()                      # [co][tap*c + ci]
    y = (cols @ wp.t()).permute(0, 2, 1).reshape(n, co, ho, wo)
    ref = F.conv2d(x, wt.half().float(), stride=2, padding=1)
    assert float((y - ref).abs().max()) < 1e-4
    dy = th.randn(n, co, ho, wo, generator=g)
    dcols = dy.reshape(n, co, ho * wo).permute(0, 2, 1) @ wp  # GEMM with the transposed pack: [n, L, 9c]
    dx = F.fold(dcols.reshape(n, ho * wo, 9, c).permute(0, 3, 2, 1).reshape(n, c * 9, ho * wo), (h, w), 3, padding=1,
                stride=2)
    xr = x.clone().requires_grad_(True)
    F.conv2d(xr, wt.half().float(), stride=2, padding=1).backward(dy)
    assert float((dx - xr.grad).abs().max()) < 1e-4


def test_checkpoint_of_the_other_label_variant_is_adopted():
    """the reference factory always builds the fork's CLIP-feature variant (script_util.py:168); a model
    built here with the default conditioning='labels' must still load such a checkpoint with strict=True (and vice
    versa), switching its label path to the checkpoint's layout."""
    from guided_diffusion_clip_b200 import script_util as su
    from oracle import golden_cfg as cfg
    fork = su.create_model(**cfg.FEAT_KW, conditioning="clip_feat")
    sd_fork = {k: v.clone() for k, v in fork.state_dict().items()}
    m = su.create_model(**cfg.FEAT_KW)  # nn.Embedding variant, 1000 classes
    assert "label_emb.weight" in m.state_dict() and not m.label_mlp
    m.load_state_dict(sd_fork, strict=True)
    assert m.label_mlp and m.num_classes == 512 and "label_emb.weight" not in m.state_dict()
    assert all(th.equal(v, sd_fork[k]) for k, v in m.state_dict().items())
    # ... and back: an upstream (Embedding) checkpoint into a fork-built model
    up = su.create_model(**cfg.FEAT_KW)
    fork.load_state_dict(up.state_dict(), strict=True)
    assert not fork.label_mlp and fork.num_classes == 1000
    # super-resolution variants differ in their forward signature: a clear error instead of a silent mismatch
    sr = su.sr_create_model(**cfg.SR_KW)
    srf = su.sr_create_model(**cfg.SR_KW, conditioning="clip_feat")
    with pytest.raises(RuntimeError, match="super-resolution variants"):
        sr.load_state_dict(srf.state_dict(), strict=True)
