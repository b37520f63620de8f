"""Tile sharding of a frame across ranks (host-side logic of the multi-GPU path; CPU-testable).

Rays shard by interleaved 32x8-pixel tiles: tile k (row-major over ceil(W/32) x ceil(H/8) tiles) belongs to rank
k % world.  Each rank holds COMPACT per-shard buffers in tile order (256 entries per tile, 8x4 sub-tiles of 32 —
the footprint of one warp fetch), every rank's buffer padded to rank 0's size so that one gather collects them.
This module mirrors `item_to_pixel` / `local_items` of csrc/traverse.cuh and csrc/tray_cuda.cu in numpy, so that a
host without the CUDA kernels (tests, a gloo run) can assemble gathered shards exactly like `tray_cuda_untile_rgba`."""
from __future__ import annotations

import numpy as np


def tiles_x(width: int) -> int:
    return (width + 31) // 32


def local_items(width: int, height: int, shard: int, shards: int) -> int:
    tiles = tiles_x(width) * ((height + 7) // 8)
    return 0 if shard >= tiles else ((tiles - shard + shards - 1) // shards) * 256


def item_pixels(width: int, height: int, shard: int, shards: int):
    """(px, py, valid) for every local work item of `shard`, in buffer order."""
    j = np.arange(local_items(width, height, shard, shards), dtype=np.int64)
    k = shard + (j >> 8) * shards
    w = j & 255
    sub, l = w >> 5, w & 31
    px = (k % tiles_x(width)) * 32 + (sub & 3) * 8 + (l & 7)
    py = (k // tiles_x(width)) * 8 + (sub >> 2) * 4 + (l >> 3)
    return px, py, (px < width) & (py < height)


def tile(frame: np.ndarray, shard: int, shards: int, pad_to: int | None = None) -> np.ndarray:
    """row-major (H, W, ...) frame -> this shard's compact buffer (invalid / padding entries are zero)."""
    h, w = frame.shape[:2]
    px, py, ok = item_pixels(w, h, shard, shards)
    n = pad_to if pad_to is not None else len(px)
    out = np.zeros((n,) + frame.shape[2:], dtype=frame.dtype)
    out[:len(px)][ok] = frame[py[ok], px[ok]]
    return out


def untile(compact: np.ndarray, frame: np.ndarray, shard: int, shards: int) -> np.ndarray:
    """scatter one shard's compact buffer into a row-major (H, W, ...) frame (in place)."""
    h, w = frame.shape[:2]
    px, py, ok = item_pixels(w, h, shard, shards)
    frame[py[ok], px[ok]] = compact[:len(px)][ok]
    return frame


def gather_frame(local_compact, width: int, height: int, rank: int, world: int, dist=None, dst: int = 0):
    """The path's one exchange step: gather every rank's compact shard on `dst` and assemble the frame there.
    `local_compact` is a torch tensor of local_items(.., 0, world) entries; returns the (H, W, ...) frame on dst."""
    import torch
    if world == 1:
        fr = np.zeros((height, width) + tuple(local_compact.shape[1:]), dtype=local_compact.numpy().dtype)
        return untile(local_compact.numpy(), fr, 0, 1)
    bufs = [torch.empty_like(local_compact) for _ in range(world)] if rank == dst else None
    dist.gather(local_compact, bufs, dst=dst)
    if rank != dst:
        return None
    fr = np.zeros((height, width) + tuple(local_compact.shape[1:]), dtype=bufs[0].numpy().dtype)
    for s in range(world):
        untile(bufs[s].numpy(), fr, s, world)
    return fr
