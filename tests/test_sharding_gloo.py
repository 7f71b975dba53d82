"""CPU, world_size 2 (gloo): the multi-GPU side of the path — per-rank seeds, label draws before the sampler,
the single all_gather, rank-major ordering and truncation — against the bookkeeping of the reference driver
(scripts/classifier_sample.py:70-102, restated in oracle_diffusion.driver_order)."""
import os
import socket

import numpy as np
import pytest
import torch as th
import torch.distributed as dist
import torch.multiprocessing as mp

from guided_diffusion_clip_b200 import dist_util
from oracle import oracle_diffusion as od

H = W = 4
NUM_CLASSES = 1000


def _fake_sampler(classes: th.Tensor, gen: th.Generator) -> th.Tensor:
    """Stands in for p_sample_loop: deterministic in (labels, this rank's generator)."""
    noise = th.randn(classes.shape[0], 3, H, W, generator=gen)
    return th.tanh(noise + classes.view(-1, 1, 1, 1).float() / NUM_CLASSES)


def _run_rank(rank, world, port, num_samples, batch, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    dist_util.setup_dist("gloo")
    gen = th.Generator().manual_seed(dist_util.rank_seed(100, rank))
    arr, labels = dist_util.sample_sharded(lambda c: _fake_sampler(c, gen), num_samples=num_samples, batch_size=batch,
                                           num_classes=NUM_CLASSES, device="cpu", base_seed=100,
                                           to_uint8=od.to_uint8_nhwc, generator=gen)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), arr, labels)
    if rank == 0:
        dist_util.save_npz(out_dir, arr, labels)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("num_samples,batch", [(10, 3), (8, 2), (5, 4)])
def test_two_rank_sharding_matches_single_process_emulation(tmp_path, num_samples, batch):
    world = 2
    mp.spawn(_run_rank, args=(world, _free_port(), num_samples, batch, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    assert np.array_equal(r0["arr_0"], r1["arr_0"]) and np.array_equal(r0["arr_1"], r1["arr_1"])  # same on all ranks
    # single-process emulation: each rank's stream replayed with its seed, stitched in the reference driver's order
    order, iters = od.driver_order(num_samples, batch, world)
    assert iters == dist_util.num_iterations(num_samples, batch, world)
    per_rank = []
    for r in range(world):
        gen = th.Generator().manual_seed(100 + r)
        its = []
        for _ in range(iters):
            classes = th.randint(0, NUM_CLASSES, (batch,), generator=gen)  # labels BEFORE the sampler's noise
            its.append((od.to_uint8_nhwc(_fake_sampler(classes, gen)).numpy(), classes.numpy()))
        per_rank.append(its)
    want_img = np.stack([per_rank[r][it][0][j] for it, r, j in order])
    want_lab = np.array([per_rank[r][it][1][j] for it, r, j in order])
    assert r0["arr_0"].dtype == np.uint8 and r0["arr_0"].shape == (num_samples, H, W, 3)
    assert np.array_equal(r0["arr_0"], want_img)       # integer outputs: bit-exact
    assert r0["arr_1"].dtype == np.int64 and np.array_equal(r0["arr_1"], want_lab)
    saved = np.load(tmp_path / f"samples_{num_samples}x{H}x{W}x3.npz")  # classifier_sample.py:103-107 naming
    assert np.array_equal(saved["arr_0"], want_img) and np.array_equal(saved["arr_1"], want_lab)


def test_iteration_count_and_slices():
    assert dist_util.num_iterations(10000, 16, 8) == 79  # ceil(10000 / 128)
    assert dist_util.num_iterations(64, 8, 8) == 1
    assert dist_util.num_iterations(65, 8, 8) == 2
    assert dist_util.shard_slice(64, 8, 3) == (24, 32)
    assert [dist_util.rank_seed(7, r) for r in range(3)] == [7, 8, 9]


def test_reference_arm_under_torchrun_prints_one_line_from_rank0():
    """The driver launches `bench.py --impl reference --gpus N` through torchrun like the GPU arm: rank 0 alone measures
    and prints the one JSON line, the other ranks exit 0 without work (CPU only, world size 2)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29547", os.path.join(root, "bench.py"),
                        "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--image-size", "64"],
                       capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
