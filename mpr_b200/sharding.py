"""Tile-row sharding of one frame across the GPUs of a box (host logic).

Top-level 64-pixel tile rows are independent (every child tile, subtape, filled pixel and
normal stays inside its parent's 64x64 screen footprint; 3D occlusion couples tiles only
along z, within one screen column), so rank r renders a contiguous band of tile rows into a
full-size image and the bands are exchanged with ONE all-gather at the end of the frame.
There is no reference counterpart: the reference is single-GPU.
"""
from __future__ import annotations

import numpy as np


def band_rows(size_px: int, world: int, rank: int) -> tuple[int, int]:
    """[row_begin, row_end) in units of 64-px tile rows.  Equal bands are required so that the
    exchange is a plain all-gather; sizes whose tile-row count is not divisible by the world
    size are rejected rather than padded."""
    rows = size_px // 64
    if rows % world != 0:
        raise ValueError(f"{rows} tile rows do not split evenly across {world} ranks")
    per = rows // world
    return rank * per, (rank + 1) * per


def band_slice(size_px: int, world: int, rank: int) -> slice:
    """Pixel rows of rank's band in a row-major (y, x) image."""
    b, e = band_rows(size_px, world, rank)
    return slice(b * 64, e * 64)


def cyclic_rows(size_px: int, world: int, rank: int) -> dict:
    """Context options for interleaved sharding: rank r owns tile rows r, r + world, r + 2*world, ...
    (geometry is rarely uniform in y, so interleaving balances far better than contiguous bands)."""
    rows = size_px // 64
    if rows % world != 0:
        raise ValueError(f"{rows} tile rows do not split evenly across {world} ranks")
    return dict(row_begin=0, row_end=rows, row_mod=world, row_rem=rank)


def all_gather_cyclic(local_full, size_px: int, group=None):
    """Interleaved counterpart of all_gather_bands: local_full (size, size) holds this rank's tile rows
    (every world-th block of 64 pixel rows).  One collective; returns the assembled image."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rows = size_px // 64
    mine = local_full.view(rows // world, world, 64, size_px)[:, rank].contiguous()      # [rows/world, 64, S]
    gathered = torch.empty((world,) + tuple(mine.shape), dtype=local_full.dtype, device=local_full.device)
    dist.all_gather_into_tensor(gathered.view(-1), mine.view(-1), group=group)
    # gathered[r, k] is tile row k * world + r
    return gathered.permute(1, 0, 2, 3).reshape(size_px, size_px)


def all_gather_bands(local_full, size_px: int, group=None):
    """local_full: torch tensor (size, size) whose band rows hold this rank's result.
    Returns the assembled (size, size) tensor on every rank; one collective."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sl = band_slice(size_px, world, rank)
    out = torch.empty_like(local_full)
    # Bands are contiguous row blocks in rank order, so the gathered buffer IS the image.
    dist.all_gather_into_tensor(out.view(-1), local_full[sl].contiguous().view(-1), group=group)
    return out
