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


def diagonal_tiles(size_px: int, world: int, rank: int) -> dict:
    """Context options for tile-cyclic sharding: rank r owns the level-0 tiles (x, y) with
    (x + y) % world == r - diagonal stripes of 64x64-px screen columns.  Much finer than whole
    rows (bear 1024^3 on 8 GPUs: 32 columns each instead of 2 rows), so the load balances even
    when the model covers a small part of the frame."""
    tiles = size_px // 64
    if tiles % world != 0:
        raise ValueError(f"{tiles} tiles per side do not split evenly across {world} ranks")
    return dict(row_begin=0, row_end=tiles, row_mod=world, row_rem=rank, col_step=1)


def diagonal_owner(size_px: int, world: int) -> np.ndarray:
    """owner[ty, tx] of every level-0 tile under diagonal_tiles()."""
    t = size_px // 64
    ty, tx = np.meshgrid(np.arange(t), np.arange(t), indexing="ij")
    return (ty + tx) % world


class TileExchange:
    """One all-gather per frame for tile-cyclic sharding.

    Every rank packs the 64x64-px blocks it owns - of ALL result images at once (depth and normals
    go out together) - into one contiguous buffer, the buffers are all-gathered, and every rank
    scatters the blocks of all ranks into full images.  The index tensors are built once per
    (size, world, device).

    `narrow`: per image, the number of bytes per pixel that actually carry information (4 = as
    is).  A 2D frame holds 0 / 1 and a depth image values below the frame size, so they travel as
    uint8 / int16 and are widened again on arrival - the assembled images are the same int32
    images, only 4x / 2x fewer bytes cross NVLink (prospero 4096^2: 64 MB -> 16 MB)."""

    def __init__(self, size_px: int, world: int, device, n_images: int = 1, narrow=None):
        import torch
        self.size, self.world, self.n_images = size_px, world, n_images
        self.narrow = list(narrow) if narrow is not None else [4] * n_images
        assert len(self.narrow) == n_images and all(b in (1, 2, 4) for b in self.narrow)
        t = size_px // 64
        owner = diagonal_owner(size_px, world)
        per = t * t // world
        self.ty, self.tx = [], []
        for r in range(world):
            ys, xs = np.nonzero(owner == r)
            assert len(ys) == per
            self.ty.append(torch.as_tensor(ys, device=device))
            self.tx.append(torch.as_tensor(xs, device=device))
        self.all_ty = torch.cat(self.ty)
        self.all_tx = torch.cat(self.tx)
        self.per = per
        block = per * 64 * 64
        self.offsets = np.concatenate([[0], np.cumsum([block * b for b in self.narrow])]).astype(np.int64)
        self.send = torch.empty(int(self.offsets[-1]), dtype=torch.uint8, device=device)
        self.recv = torch.empty((world, int(self.offsets[-1])), dtype=torch.uint8, device=device)
        self._dtype = {1: torch.uint8, 2: torch.int16, 4: torch.int32}

    def _blocks(self, image):
        t = self.size // 64
        return image.view(t, 64, t, 64).permute(0, 2, 1, 3)          # [ty, tx, y, x] view, no copy

    def _segment(self, buf, k):
        """View of image k's slot in a packed buffer (last dim = bytes) as [.., per, 64, 64]."""
        seg = buf[..., int(self.offsets[k]):int(self.offsets[k + 1])]
        return seg.view(self._dtype[self.narrow[k]]).view(*buf.shape[:-1], self.per, 64, 64)

    def pack(self, images, rank: int):
        for k, img in enumerate(images):
            self._segment(self.send, k).copy_(self._blocks(img)[self.ty[rank], self.tx[rank]])
        return self.send

    def unpack(self, outs):
        for k, out in enumerate(outs):
            blocks = self._segment(self.recv, k).reshape(-1, 64, 64)
            self._blocks(out)[self.all_ty, self.all_tx] = blocks.to(out.dtype)
        return outs

    def gather(self, images, outs, group=None):
        """images: this rank's full-size int32 tensors (only its tiles are meaningful);
        outs: full-size tensors that receive the assembled frame."""
        import torch.distributed as dist
        self.pack(images, dist.get_rank(group))
        dist.all_gather_into_tensor(self.recv.view(-1), self.send, group=group)
        return self.unpack(outs)


class NativeExchange:
    """TileExchange with the library's own pack / unpack kernels (mprb_exchange_*): one launch to
    narrow and gather the owned blocks, one NCCL all-gather, one launch to scatter all ranks' blocks
    into the context's full-size image and normals - in place, on the caller's CUDA stream."""

    def __init__(self, ctx, dim: int, world: int, device):
        import torch
        self.ctx, self.dim, self.world = ctx, dim, world
        self.bytes = ctx.exchange_bytes(dim)
        if self.bytes == 0:
            raise ValueError("context is not sharded over the whole frame")
        self.send = torch.empty(self.bytes, dtype=torch.uint8, device=device)
        self.recv = torch.empty(world * self.bytes, dtype=torch.uint8, device=device)

    def gather(self, group=None):
        import torch
        import torch.distributed as dist
        stream = torch.cuda.current_stream().cuda_stream
        self.ctx.exchange_pack(self.dim, self.send.data_ptr(), stream)
        dist.all_gather_into_tensor(self.recv, self.send, group=group)
        self.ctx.exchange_unpack(self.dim, self.recv.data_ptr(), stream)
