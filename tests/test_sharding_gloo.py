"""N > 1 host logic on CPU: world_size-2 gloo processes shard a frame by interleaved tiles, fill their compact film
with a known function of the pixel, all-gather once, and reassemble the raster (what bench.py does over NCCL with
rtb_film_device_buffer / rtb_unshard_film)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from raytracer_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, w, h, tile, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lists, slots = sharding.shard_pixel_lists(w, h, world, tile)
    mine = lists[rank]
    film = torch.zeros(slots, 3)
    pix = torch.from_numpy(mine.astype(np.int64))
    film[: len(mine), 0] = (pix % w).float()
    film[: len(mine), 1] = (pix // w).float()
    film[: len(mine), 2] = float(rank + 1)
    gathered = torch.empty(world * slots, 3)
    dist.all_gather_into_tensor(gathered, film)
    counts = torch.tensor([float(len(mine))])
    dist.all_reduce(counts)
    if rank == 0:
        img = sharding.unshard(gathered.numpy(), w, h, world, tile)
        np.save(out_path, img)
        assert int(counts[0]) == w * h
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("w,h,tile", [(200, 152, 32), (64, 64, 16), (33, 7, 32)])
def test_two_rank_gather_reassembles_the_frame(tmp_path, w, h, tile):
    out = str(tmp_path / "img.npy")
    mp.spawn(_worker, args=(2, _free_port(), w, h, tile, out), nprocs=2, join=True)
    img = np.load(out)
    ys, xs = np.mgrid[0:h, 0:w]
    assert (img[..., 0] == xs).all() and (img[..., 1] == ys).all()
    owner = img[..., 2]
    assert set(np.unique(owner)) <= {1.0, 2.0}
    tiles_x = (w + tile - 1) // tile
    expect = ((ys // tile) * tiles_x + (xs // tile)) % 2 + 1
    assert (owner == expect).all()


def test_shard_lists_partition_every_pixel_once():
    for n in (1, 2, 3, 4, 8):
        lists, slots = sharding.shard_pixel_lists(1920, 1080, n)
        allpix = np.concatenate(lists)
        assert len(allpix) == 1920 * 1080 and len(np.unique(allpix)) == 1920 * 1080
        assert all(len(l) <= slots for l in lists)
        # interleaving keeps the shards balanced to within a few tiles
        assert max(len(l) for l in lists) - min(len(l) for l in lists) <= 4 * 32 * 32
