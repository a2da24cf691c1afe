"""N>1 host path on CPU: band partition + single all-gather over gloo, world_size 2."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, load_tape
from mpr_b200 import sharding


def test_band_rows_partition_the_image():
    for size, world in [(1024, 2), (1024, 8), (4096, 8), (2048, 4)]:
        rows = [sharding.band_rows(size, world, r) for r in range(world)]
        assert rows[0][0] == 0 and rows[-1][1] == size // 64
        assert all(rows[i][1] == rows[i + 1][0] for i in range(world - 1))
    with pytest.raises(ValueError):
        sharding.band_rows(1024 + 64, 2, 0)    # 17 rows do not split in two


def test_diagonal_tiles_partition_the_frame():
    for size, world in [(1024, 2), (1024, 8), (4096, 8), (2048, 4), (256, 2)]:
        owner = sharding.diagonal_owner(size, world)
        counts = np.bincount(owner.ravel(), minlength=world)
        assert (counts == (size // 64) ** 2 // world).all()          # equal shares: a plain all-gather works
        for r in range(world):
            o = sharding.diagonal_tiles(size, world, r)
            assert o["col_step"] == 1 and o["row_mod"] == world and o["row_rem"] == r
        # every tile row and every tile column is spread over all ranks
        if size // 64 >= world:
            assert all(len(set(owner[k])) == world and len(set(owner[:, k])) == world for k in range(size // 64))
    with pytest.raises(ValueError):
        sharding.diagonal_tiles(1024 + 64, 2, 0)


def _worker(rank, world, port, size, full_path, out_path):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.from_numpy(np.load(full_path))
    local = torch.zeros_like(full)
    sl = sharding.band_slice(size, world, rank)
    local[sl] = full[sl]                        # this rank only "rendered" its band
    got = sharding.all_gather_bands(local, size)
    ok = bool(torch.equal(got, full))
    owner = (torch.arange(size) // 64) % world
    local2 = torch.zeros_like(full)
    local2[owner == rank] = full[owner == rank]            # interleaved assignment
    ok = ok and bool(torch.equal(sharding.all_gather_cyclic(local2, size), full))
    # tile-cyclic (diagonal) assignment, two images in one exchange
    own = torch.from_numpy(np.kron(sharding.diagonal_owner(size, world) == rank, np.ones((64, 64), dtype=bool)))
    second = full * 3 + 1
    a, b = torch.where(own, full, torch.full_like(full, -7)), torch.where(own, second, torch.full_like(full, -7))
    ex = sharding.TileExchange(size, world, "cpu", n_images=2)
    out_a, out_b = torch.empty_like(full), torch.empty_like(full)
    ex.gather([a, b], [out_a, out_b])
    ok = ok and bool(torch.equal(out_a, full)) and bool(torch.equal(out_b, second))
    # narrow transport: a 0/1 image as bytes, a depth-like image as int16, a full-width one beside them
    depth = (full * 517 + 3) % 1024
    wide = full * 0x01010101 - 5
    ex = sharding.TileExchange(size, world, "cpu", n_images=3, narrow=[1, 2, 4])
    outs = [torch.empty_like(full) for _ in range(3)]
    ex.gather([torch.where(own, x, torch.zeros_like(x)) for x in (full, depth, wide)], outs)
    ok = ok and all(bool(torch.equal(o, x)) for o, x in zip(outs, (full, depth, wide)))
    np.save(out_path.format(rank), np.array([ok]))
    dist.destroy_process_group()


def test_all_gather_reassembles_the_frame_gloo(tmp_path):
    import oracle
    import torch.multiprocessing as mp
    size = 256
    o = oracle.CpuOracle(size)
    o.render2D(load_tape("hello_world"))
    full = np.array(o.image(), dtype=np.int32)
    o.close()
    fp = tmp_path / "full.npy"
    np.save(fp, full)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "ok{}.npy")
    mp.spawn(_worker, args=(2, port, size, str(fp), out), nprocs=2, join=True)
    assert all(bool(np.load(out.format(r))[0]) for r in range(2))
