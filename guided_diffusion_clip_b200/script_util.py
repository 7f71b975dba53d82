"""Factories with the reference's names, keyword arguments and defaults
(guided_diffusion/script_util.py:12-477) returning this package's CUDA-backed objects.

One documented divergence: the fork's `create_model` unconditionally returns
`UNetModel_clip_feat` with NUM_CLASSES=512.  BASELINE's "class-cond ADM" configs are the upstream
semantics (nn.Embedding over 1000 classes fed by int labels), and the fork's own classifier_sample.py
cannot run against its own factory.  `create_model(..., conditioning=...)` therefore selects:
  "labels" (default)  -> unet.UNetModel,         num_classes = NUM_LABEL_CLASSES (1000), model(x, t, y=int64[N])
  "clip_feat"         -> unet.UNetModel_clip_feat, num_classes = NUM_CLASSES (512),      model(x, t, clip_feat=...)
Everything else — argument names, defaults dictionaries, channel_mult tables, attention_ds — is the reference's.
"""
from __future__ import annotations

import argparse
import inspect

import yaml

from . import gaussian_diffusion as gd
from .respace import SpacedDiffusion, space_timesteps
from .unet import EncoderUNetModel, SRImageModel_Feat, SuperResModel, UNetModel, UNetModel_clip_feat

NUM_CLASSES = 512  # fork constant: width of the CLIP feature (script_util.py:9)
NUM_LABEL_CLASSES = 1000  # upstream ImageNet label count == classifier out_channels (script_util.py:260)

_CHANNEL_MULT = {512: (0.5, 1, 1, 2, 2, 4, 4), 256: (1, 1, 2, 2, 4, 4), 128: (1, 1, 2, 3, 4), 64: (1, 2, 3, 4)}
_SR_CHANNEL_MULT = {512: (1, 1, 2, 2, 4, 4), 256: (1, 1, 2, 2, 4, 4), 128: (1, 1, 2, 3, 4), 64: (1, 2, 3, 4)}


def diffusion_defaults():
    """script_util.py:12-25."""
    return dict(learn_sigma=False, diffusion_steps=1000, noise_schedule="linear", timestep_respacing="",
                use_kl=False, predict_xstart=False, rescale_timesteps=False, rescale_learned_sigmas=False)


def classifier_defaults():
    """script_util.py:28-41."""
    return dict(image_size=64, classifier_use_fp16=False, classifier_width=128, classifier_depth=2,
                classifier_attention_resolutions="32,16,8", classifier_use_scale_shift_norm=True,
                classifier_resblock_updown=True, classifier_pool="attention")


def model_and_diffusion_defaults():
    """script_util.py:44-66."""
    res = dict(image_size=64, num_channels=128, num_res_blocks=2, num_heads=4, num_heads_upsample=-1,
               num_head_channels=-1, attention_resolutions="16,8", channel_mult="", dropout=0.0, class_cond=False,
               use_checkpoint=False, use_scale_shift_norm=True, resblock_updown=False, use_fp16=False,
               use_new_attention_order=False)
    res.update(diffusion_defaults())
    return res


def classifier_and_diffusion_defaults():
    res = classifier_defaults()
    res.update(diffusion_defaults())
    return res


def _attention_ds(image_size: int, resolutions: str):
    return tuple(image_size // int(r) for r in resolutions.split(","))


def create_model_and_diffusion(image_size, class_cond, learn_sigma, num_channels, num_res_blocks, channel_mult,
                               num_heads, num_head_channels, num_heads_upsample, attention_resolutions, dropout,
                               diffusion_steps, noise_schedule, timestep_respacing, use_kl, predict_xstart,
                               rescale_timesteps, rescale_learned_sigmas, use_checkpoint, use_scale_shift_norm,
                               resblock_updown, use_fp16, use_new_attention_order, conditioning="labels"):
    """script_util.py:75-128."""
    model = create_model(
        image_size, num_channels, num_res_blocks, channel_mult=channel_mult, learn_sigma=learn_sigma,
        class_cond=class_cond, use_checkpoint=use_checkpoint, attention_resolutions=attention_resolutions,
        num_heads=num_heads, num_head_channels=num_head_channels, num_heads_upsample=num_heads_upsample,
        use_scale_shift_norm=use_scale_shift_norm, dropout=dropout, resblock_updown=resblock_updown,
        use_fp16=use_fp16, use_new_attention_order=use_new_attention_order, conditioning=conditioning)
    diffusion = create_gaussian_diffusion(
        steps=diffusion_steps, learn_sigma=learn_sigma, noise_schedule=noise_schedule, use_kl=use_kl,
        predict_xstart=predict_xstart, rescale_timesteps=rescale_timesteps,
        rescale_learned_sigmas=rescale_learned_sigmas, timestep_respacing=timestep_respacing)
    return model, diffusion


def create_model(image_size, num_channels, num_res_blocks, channel_mult="", learn_sigma=False, class_cond=False,
                 use_checkpoint=False, attention_resolutions="16", num_heads=1, num_head_channels=-1,
                 num_heads_upsample=-1, use_scale_shift_norm=False, dropout=0, resblock_updown=False, use_fp16=False,
                 use_new_attention_order=False, conditioning="labels"):
    """script_util.py:131-187 (channel_mult table :149-161, attention_ds :163-165)."""
    if channel_mult == "":
        if image_size not in _CHANNEL_MULT:
            raise ValueError(f"unsupported image size: {image_size}")
        channel_mult = _CHANNEL_MULT[image_size]
    else:
        channel_mult = tuple(int(m) for m in channel_mult.split(","))
    if conditioning == "labels":
        cls, n_cls = UNetModel, NUM_LABEL_CLASSES
    elif conditioning == "clip_feat":
        cls, n_cls = UNetModel_clip_feat, NUM_CLASSES
    else:
        raise ValueError(f"unknown conditioning {conditioning!r}")
    return cls(
        image_size=image_size, in_channels=3, model_channels=num_channels,
        out_channels=(3 if not learn_sigma else 6), num_res_blocks=num_res_blocks,
        attention_resolutions=_attention_ds(image_size, attention_resolutions), dropout=dropout,
        channel_mult=channel_mult, num_classes=(n_cls if class_cond else None), use_checkpoint=use_checkpoint,
        use_fp16=use_fp16, num_heads=num_heads, num_head_channels=num_head_channels,
        num_heads_upsample=num_heads_upsample, use_scale_shift_norm=use_scale_shift_norm,
        resblock_updown=resblock_updown, use_new_attention_order=use_new_attention_order)


def create_classifier_and_diffusion(image_size, classifier_use_fp16, classifier_width, classifier_depth,
                                    classifier_attention_resolutions, classifier_use_scale_shift_norm,
                                    classifier_resblock_updown, classifier_pool, learn_sigma, diffusion_steps,
                                    noise_schedule, timestep_respacing, use_kl, predict_xstart, rescale_timesteps,
                                    rescale_learned_sigmas):
    """script_util.py:190-228."""
    classifier = create_classifier(image_size, classifier_use_fp16, classifier_width, classifier_depth,
                                   classifier_attention_resolutions, classifier_use_scale_shift_norm,
                                   classifier_resblock_updown, classifier_pool)
    diffusion = create_gaussian_diffusion(
        steps=diffusion_steps, learn_sigma=learn_sigma, noise_schedule=noise_schedule, use_kl=use_kl,
        predict_xstart=predict_xstart, rescale_timesteps=rescale_timesteps,
        rescale_learned_sigmas=rescale_learned_sigmas, timestep_respacing=timestep_respacing)
    return classifier, diffusion


def create_classifier(image_size, classifier_use_fp16, classifier_width, classifier_depth,
                      classifier_attention_resolutions, classifier_use_scale_shift_norm, classifier_resblock_updown,
                      classifier_pool):
    """script_util.py:231-269: out_channels=1000 and num_head_channels=64 are hard-coded there too."""
    if image_size not in _CHANNEL_MULT:
        raise ValueError(f"unsupported image size: {image_size}")
    return EncoderUNetModel(
        image_size=image_size, in_channels=3, model_channels=classifier_width, out_channels=1000,
        num_res_blocks=classifier_depth,
        attention_resolutions=_attention_ds(image_size, classifier_attention_resolutions),
        channel_mult=_CHANNEL_MULT[image_size], use_fp16=classifier_use_fp16, num_head_channels=64,
        use_scale_shift_norm=classifier_use_scale_shift_norm, resblock_updown=classifier_resblock_updown,
        pool=classifier_pool)


def sr_model_and_diffusion_defaults():
    """script_util.py:272-280 (fork values large_size=128, small_size=64)."""
    res = model_and_diffusion_defaults()
    res["large_size"] = 128
    res["small_size"] = 64
    arg_names = inspect.getfullargspec(sr_create_model_and_diffusion)[0]
    for k in list(res.keys()):
        if k not in arg_names:
            del res[k]
    return res


def sr_create_model_and_diffusion(large_size, small_size, class_cond, learn_sigma, num_channels, num_res_blocks,
                                  num_heads, num_head_channels, num_heads_upsample, attention_resolutions, dropout,
                                  diffusion_steps, noise_schedule, timestep_respacing, use_kl, predict_xstart,
                                  rescale_timesteps, rescale_learned_sigmas, use_checkpoint, use_scale_shift_norm,
                                  resblock_updown, use_fp16):
    """script_util.py:283-332."""
    model = sr_create_model(large_size, small_size, num_channels, num_res_blocks, learn_sigma=learn_sigma,
                            class_cond=class_cond, use_checkpoint=use_checkpoint,
                            attention_resolutions=attention_resolutions, num_heads=num_heads,
                            num_head_channels=num_head_channels, num_heads_upsample=num_heads_upsample,
                            use_scale_shift_norm=use_scale_shift_norm, dropout=dropout,
                            resblock_updown=resblock_updown, use_fp16=use_fp16)
    diffusion = create_gaussian_diffusion(
        steps=diffusion_steps, learn_sigma=learn_sigma, noise_schedule=noise_schedule, use_kl=use_kl,
        predict_xstart=predict_xstart, rescale_timesteps=rescale_timesteps,
        rescale_learned_sigmas=rescale_learned_sigmas, timestep_respacing=timestep_respacing)
    return model, diffusion


def sr_create_model(large_size, small_size, num_channels, num_res_blocks, learn_sigma, class_cond, use_checkpoint,
                    attention_resolutions, num_heads, num_head_channels, num_heads_upsample, use_scale_shift_norm,
                    dropout, resblock_updown, use_fp16, conditioning="low_res"):
    """script_util.py:335-389.  conditioning="low_res" -> unet.SuperResModel (upstream, BASELINE config 4);
    "clip_feat" -> the fork's SRImageModel_Feat (script_util.py:371)."""
    _ = small_size
    if large_size not in _SR_CHANNEL_MULT:
        raise ValueError(f"unsupported large size: {large_size}")
    if conditioning == "low_res":
        cls, n_cls = SuperResModel, NUM_LABEL_CLASSES
    elif conditioning == "clip_feat":
        cls, n_cls = SRImageModel_Feat, NUM_CLASSES
    else:
        raise ValueError(f"unknown conditioning {conditioning!r}")
    return cls(
        image_size=large_size, in_channels=3, model_channels=num_channels,
        out_channels=(3 if not learn_sigma else 6), num_res_blocks=num_res_blocks,
        attention_resolutions=_attention_ds(large_size, attention_resolutions), dropout=dropout,
        channel_mult=_SR_CHANNEL_MULT[large_size], num_classes=(n_cls if class_cond else None),
        use_checkpoint=use_checkpoint, num_heads=num_heads, num_head_channels=num_head_channels,
        num_heads_upsample=num_heads_upsample, use_scale_shift_norm=use_scale_shift_norm,
        resblock_updown=resblock_updown, use_fp16=use_fp16)


def create_gaussian_diffusion(*, steps=1000, learn_sigma=False, sigma_small=False, noise_schedule="linear",
                              use_kl=False, predict_xstart=False, rescale_timesteps=False,
                              rescale_learned_sigmas=False, timestep_respacing=""):
    """script_util.py:392-430."""
    betas = gd.get_named_beta_schedule(noise_schedule, steps)
    if use_kl:
        loss_type = gd.LossType.RESCALED_KL
    elif rescale_learned_sigmas:
        loss_type = gd.LossType.RESCALED_MSE
    else:
        loss_type = gd.LossType.MSE
    if not timestep_respacing:
        timestep_respacing = [steps]
    if learn_sigma:
        var_type = gd.ModelVarType.LEARNED_RANGE
    else:
        var_type = gd.ModelVarType.FIXED_SMALL if sigma_small else gd.ModelVarType.FIXED_LARGE
    return SpacedDiffusion(
        use_timesteps=space_timesteps(steps, timestep_respacing), betas=betas,
        model_mean_type=(gd.ModelMeanType.START_X if predict_xstart else gd.ModelMeanType.EPSILON),
        model_var_type=var_type, loss_type=loss_type, rescale_timesteps=rescale_timesteps)


# ---- argparse / YAML helpers (script_util.py:433-477) ----------------------------------------------
def add_dict_to_argparser(parser, default_dict):
    for k, v in default_dict.items():
        v_type = type(v)
        if v is None:
            v_type = str
        elif isinstance(v, bool):
            v_type = str2bool
        parser.add_argument(f"--{k}", default=v, type=v_type)
    parser.add_argument("--config-file", dest="config_file", default=None, type=argparse.FileType(mode="r"))
    parser.add_argument("-d", "--description", dest="description", type=str, default="",
                        help="free description of the run")


def args_to_dict(args, keys):
    return {k: getattr(args, k) for k in keys}


def str2bool(v):
    if isinstance(v, bool):
        return v
    if v.lower() in ("yes", "true", "t", "y", "1"):
        return True
    if v.lower() in ("no", "false", "f", "n", "0"):
        return False
    raise argparse.ArgumentTypeError("boolean value expected")


def parse_yaml(args):
    if getattr(args, "config_file", None):
        data = yaml.load(args.config_file, yaml.SafeLoader)
        delattr(args, "config_file")
        arg_dict = args.__dict__
        for key, value in data.items():
            if isinstance(value, list):
                for v in value:
                    arg_dict[key].append(v)
            else:
                arg_dict[key] = value
    return args
