"""guided_diffusion_clip_b200 — B200-native (sm_100a) drop-in for the sampling hot path of
ErezYosef/guided-diffusion-clip: same factories, same sampling API, same state_dict layout; every hot
operation is a hand-written CUDA kernel reached through the C ABI in include/gd_b200.h."""
from . import _lib  # noqa: F401  (does not load the .so until first use)

__all__ = ["script_util", "gaussian_diffusion", "respace", "unet", "dist_util", "sampler", "engine"]
