"""CPU tests of the multi-GPU host logic: the tile-shard mapping (numpy mirror of the device mapping) and the
gather-to-rank-0 exchange step under torch.distributed with the gloo backend, world_size 2 and 3."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tray_racing_b200 import cuda, sharding


@pytest.mark.parametrize("w,h", [(1920, 1080), (100, 37), (33, 9), (64, 16)])
@pytest.mark.parametrize("shards", [1, 2, 3, 8])
def test_shards_partition_every_pixel_once(w, h, shards):
    seen = np.zeros((h, w), dtype=np.int32)
    for s in range(shards):
        px, py, ok = sharding.item_pixels(w, h, s, shards)
        assert len(px) == sharding.local_items(w, h, s, shards) == cuda.local_items(w, h, s, shards)
        assert len(px) <= sharding.local_items(w, h, 0, shards)             # rank 0's buffer is the largest
        np.add.at(seen, (py[ok], px[ok]), 1)
        assert ok.sum() == cuda.shard_pixels(w, h, s, shards)               # agrees with the C ABI's count
        assert (cuda.shard_mask(w, h, s, shards).reshape(h, w)[py[ok], px[ok]]).all()
    assert (seen == 1).all()


def test_warp_fetch_footprint_is_an_8x4_pixel_block():
    px, py, ok = sharding.item_pixels(1920, 1080, 0, 1)
    for j0 in (0, 32, 256, 256 * 61 + 96):
        x, y = px[j0:j0 + 32], py[j0:j0 + 32]
        assert x.max() - x.min() == 7 and y.max() - y.min() == 3


def test_tile_untile_roundtrip():
    rng = np.random.default_rng(0)
    fr = rng.integers(0, 255, size=(37, 100, 4), dtype=np.uint8)
    for shards in (1, 2, 5):
        out = np.zeros_like(fr)
        pad = sharding.local_items(100, 37, 0, shards)
        for s in range(shards):
            sharding.untile(sharding.tile(fr, s, shards, pad_to=pad), out, s, shards)
        assert (out == fr).all()


def _worker(rank, world, port, w, h, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(42)                                     # every rank knows the "true" frame
        truth = rng.integers(0, 2 ** 31 - 1, size=(h, w), dtype=np.int32)
        pad = sharding.local_items(w, h, 0, world)
        local = torch.from_numpy(sharding.tile(truth, rank, world, pad_to=pad))   # what this rank's GPU would hold
        frame = sharding.gather_frame(local, w, h, rank, world, dist)
        # weak-scaling bookkeeping of bench.py: rays summed, time max-reduced
        rays = torch.tensor([float(cuda.shard_pixels(w, h, rank, world))], dtype=torch.float64)
        ms = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(rays, op=dist.ReduceOp.SUM)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            q.put((bool((frame == truth).all()), float(rays.item()), float(ms.item())))
        else:
            assert frame is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,w,h", [(2, 200, 120), (3, 101, 37)])
def test_gather_to_rank0_over_gloo(world, w, h):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, w, h, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, rays, ms = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and rays == w * h and ms == float(world)
