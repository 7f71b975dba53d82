"""Batch sharding across the GPUs of one box: the role of guided_diffusion/dist_util.py:21-83 plus the
sampling-driver bookkeeping of scripts/classifier_sample.py:70-107, without MPI.

One process per GPU (torchrun-style env rendezvous), every rank samples its own batch with seed
`base_seed + rank`; the ONLY collective is the final all_gather of uint8 NHWC samples and int64 labels
(NCCL over NVLink on GPUs, gloo in the CPU tests).  The integer bookkeeping — iteration count, rank-major
output order, truncation to num_samples — follows the reference driver exactly.
"""
from __future__ import annotations

import ctypes as C
import io
import os
from typing import Callable, List, Optional, Tuple

import numpy as np
import torch as th
import torch.distributed as dist

from . import _lib as L


def setup_dist(backend: Optional[str] = None) -> None:
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (dist_util.py:21-42 without mpi4py) and
    bind this process to cuda:LOCAL_RANK (the reference leaves that to the launcher, dist_util.py:27,49-50)."""
    if dist.is_initialized():
        return
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    if backend is None:
        backend = "nccl" if th.cuda.is_available() else "gloo"
    if th.cuda.is_available():
        th.cuda.set_device(local % th.cuda.device_count())
    dist.init_process_group(backend=backend, rank=rank, world_size=world)


def dev() -> th.device:
    """dist_util.py:45-51."""
    if th.cuda.is_available():
        return th.device("cuda", th.cuda.current_device())
    return th.device("cpu")


def world_size() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_initialized() else 0


def load_state_dict(path, **kwargs):
    """dist_util.py:54-74 semantics (a flat OrderedDict[str, Tensor]) without the MPI broadcast: every rank of a
    single box reads the file itself."""
    with open(path, "rb") as f:
        data = f.read()
    return th.load(io.BytesIO(data), **kwargs)


# ---- sharding bookkeeping (bit-exact with scripts/classifier_sample.py:70,93,99-102) -------------------
def num_iterations(num_samples: int, batch_size: int, world: int) -> int:
    """`while len(all_images) * batch_size < num_samples` with `world` entries appended per iteration."""
    it, n_lists = 0, 0
    while n_lists * batch_size < num_samples:
        n_lists += world
        it += 1
    return it


def rank_seed(base_seed: int, r: int) -> int:
    """The reference never seeds; this package defines seed = base + rank."""
    return int(base_seed) + int(r)


def shard_slice(global_batch: int, world: int, r: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) share of a global batch for rank r (global_batch must divide evenly)."""
    assert global_batch % world == 0, "global batch must be divisible by the world size"
    per = global_batch // world
    return r * per, (r + 1) * per


def to_uint8_nhwc(sample: th.Tensor) -> th.Tensor:
    """((sample + 1) * 127.5).clamp(0, 255).to(uint8).permute(0, 2, 3, 1) (classifier_sample.py:87-89)."""
    if sample.device.type != "cuda":
        raise L.GdError("to_uint8_nhwc only runs on CUDA; there is no CPU path")
    n, c, h, w = sample.shape
    x = sample.float().contiguous()
    out = th.empty((n, h, w, c), dtype=th.uint8, device=sample.device)
    with th.cuda.device(sample.device):
        stream = C.c_void_p(th.cuda.current_stream().cuda_stream)
        L.check(L.load().gd_to_uint8_nhwc(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), n, c, h, w, stream),
                "gd_to_uint8_nhwc")
    return out


def all_gather_batch(sample_u8: th.Tensor, labels: th.Tensor) -> Tuple[List[th.Tensor], List[th.Tensor]]:
    """The one collective of the path (classifier_sample.py:91-96): per-rank lists in rank order.  One
    all_gather_into_tensor per tensor into a single [world * B, ...] buffer (no per-rank zeros_like lists)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [sample_u8], [labels]
    w = dist.get_world_size()
    imgs = th.empty((w * sample_u8.shape[0],) + tuple(sample_u8.shape[1:]), dtype=sample_u8.dtype, device=sample_u8.device)
    labs = th.empty((w * labels.shape[0],) + tuple(labels.shape[1:]), dtype=labels.dtype, device=labels.device)
    dist.all_gather_into_tensor(imgs, sample_u8.contiguous())
    dist.all_gather_into_tensor(labs, labels.contiguous())
    return list(imgs.chunk(w, 0)), list(labs.chunk(w, 0))


class GatherBuffer:
    """Output stage of the sampling driver (classifier_sample.py:87-96) without intermediate copies: ONE preallocated
    [world * B, H, W, C] uint8 buffer (and [world * B] int64 labels); the uint8-NHWC pack kernel of rank r writes its
    finished batch straight into rows [r * B, (r + 1) * B) and one in-place all_gather_into_tensor per tensor (NCCL
    over NVLink) fills in the other ranks' rows.  Bit-identical to to_uint8_nhwc + all_gather_batch."""

    def __init__(self, batch: int, c: int, h: int, w: int, device):
        self.world, self.rank, self.batch = world_size(), rank(), batch
        self.images = th.empty((self.world * batch, h, w, c), dtype=th.uint8, device=device)
        self.labels = th.empty((self.world * batch,), dtype=th.int64, device=device)

    def pack_and_gather(self, sample: th.Tensor, labels: th.Tensor) -> Tuple[List[th.Tensor], List[th.Tensor]]:
        if sample.device.type != "cuda":
            raise L.GdError("GatherBuffer only runs on CUDA; there is no CPU path")
        n, c, h, w = sample.shape
        assert n == self.batch and (h, w, c) == tuple(self.images.shape[1:]), (sample.shape, self.images.shape)
        lo, hi = self.rank * n, (self.rank + 1) * n
        x = sample.float().contiguous()
        mine = self.images[lo:hi]
        with th.cuda.device(sample.device):
            stream = C.c_void_p(th.cuda.current_stream().cuda_stream)
            L.check(L.load().gd_to_uint8_nhwc(C.c_void_p(x.data_ptr()), C.c_void_p(mine.data_ptr()), n, c, h, w, stream),
                    "gd_to_uint8_nhwc")
        self.labels[lo:hi].copy_(labels)
        if self.world > 1:
            dist.all_gather_into_tensor(self.images, mine)          # in place: `mine` is this rank's slice
            dist.all_gather_into_tensor(self.labels, self.labels[lo:hi])
        return list(self.images.chunk(self.world, 0)), list(self.labels.chunk(self.world, 0))


def sample_sharded(sample_batch: Callable[[th.Tensor], th.Tensor], *, num_samples: int, batch_size: int,
                   num_classes: int, device, base_seed: int = 0, to_uint8: Callable = to_uint8_nhwc,
                   generator: Optional[th.Generator] = None):
    """The sampling driver loop of scripts/classifier_sample.py:70-102.

    `sample_batch(classes) -> [batch_size, 3, H, W] float` runs one full sampling loop on this rank.
    Labels are drawn per iteration BEFORE the sampler's noise (classifier_sample.py:72-74) from this rank's
    generator.  Returns (arr uint8 [num_samples,H,W,3], label_arr int64 [num_samples]) — identical on all ranks."""
    all_images: List[np.ndarray] = []
    all_labels: List[np.ndarray] = []
    gbuf: Optional[GatherBuffer] = None
    while len(all_images) * batch_size < num_samples:
        classes = th.randint(low=0, high=num_classes, size=(batch_size,), device=device, generator=generator)
        sample = sample_batch(classes)
        if to_uint8 is to_uint8_nhwc and sample.device.type == "cuda":
            # fused output stage: the pack kernel writes into this rank's rows of the gather buffer
            if gbuf is None:
                gbuf = GatherBuffer(batch_size, sample.shape[1], sample.shape[2], sample.shape[3], sample.device)
            imgs, labs = gbuf.pack_and_gather(sample, classes)
        else:
            u8 = to_uint8(sample).contiguous()
            imgs, labs = all_gather_batch(u8, classes)
        all_images.extend(t.cpu().numpy() for t in imgs)
        all_labels.extend(t.cpu().numpy() for t in labs)
    arr = np.concatenate(all_images, axis=0)[:num_samples]
    label_arr = np.concatenate(all_labels, axis=0)[:num_samples]
    return arr, label_arr


def save_npz(out_dir: str, arr: np.ndarray, label_arr: Optional[np.ndarray] = None) -> Optional[str]:
    """Rank 0 writes samples_{N}x{H}x{W}x3.npz with arr_0 / arr_1 (classifier_sample.py:103-107)."""
    if rank() != 0:
        return None
    shape_str = "x".join(str(x) for x in arr.shape)
    path = os.path.join(out_dir, f"samples_{shape_str}.npz")
    if label_arr is None:
        np.savez(path, arr)
    else:
        np.savez(path, arr, label_arr)
    return path


def load_data_for_worker(base_samples: str, batch_size: int, class_cond: bool, rank_: Optional[int] = None,
                         world: Optional[int] = None):
    """Low-resolution conditioning stream of the upsampler (scripts/super_res_sample.py:77-100).

    Reads the base sampler's npz (`arr_0` uint8 NHWC, `arr_1` labels) and yields, forever, dicts with
    `low_res` float32 NCHW in [-1, 1] (`uint8 / 127.5 - 1`) and, if class-conditional, `y` int64.  Rank r takes the
    base samples r, r+W, r+2W, ... and starts over at r when the file is exhausted; a partially filled batch is
    carried into the next pass (the reference never resets its buffer), so batch k of rank r is the same on every
    run.  The tensors are CPU tensors, exactly as in the reference; move them with `.to(dist_util.dev())`."""
    with np.load(base_samples) as obj:
        image_arr = obj["arr_0"]
        label_arr = obj["arr_1"] if class_cond else None
    r = rank() if rank_ is None else rank_
    w = world_size() if world is None else world
    if len(image_arr) <= r:
        raise ValueError(f"base_samples holds {len(image_arr)} images, none for rank {r} of {w}")
    images: List[np.ndarray] = []
    labels: List[np.ndarray] = []
    while True:
        for i in range(r, len(image_arr), w):
            images.append(image_arr[i])
            if class_cond:
                labels.append(label_arr[i])
            if len(images) < batch_size:
                continue
            batch = th.from_numpy(np.stack(images)).float() / 127.5 - 1.0
            out = {"low_res": batch.permute(0, 3, 1, 2)}
            if class_cond:
                out["y"] = th.from_numpy(np.stack(labels))
            yield out
            images, labels = [], []
