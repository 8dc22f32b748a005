"""CPU (gloo, world_size 2 and 3): the multi-GPU sharding logic of bench.py -- contiguous tile-row bands per rank and one
gather of RGB bands to rank 0 -- reproduces the single-process image."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libheif_b200 import sharding  # noqa: E402


def test_band_partition_covers_all_tiles():
    for rows, cols, world in [(16, 16, 1), (16, 16, 2), (16, 16, 8), (5, 3, 2), (7, 4, 3), (2, 2, 4)]:
        seen = []
        for r in range(world):
            r0, n, idx = sharding.my_band(rows, cols, world, r)
            assert idx == [rr * cols + c for rr in range(r0, r0 + n) for c in range(cols)]
            seen += idx
        assert seen == list(range(rows * cols))


def _worker(rank, world, port, tile_rows, tile_h, row_bytes, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full_ref = torch.from_numpy(np.random.RandomState(7).randint(0, 256, (tile_rows * tile_h, row_bytes)).astype(np.uint8))
    r0, n, _ = sharding.my_band(tile_rows, 1, world, rank)
    band = full_ref[r0 * tile_h:(r0 + n) * tile_h].clone()      # what this rank's GPU would have produced
    got = sharding.gather_bands(band, tile_rows, tile_h, world, rank)
    if rank == 0:
        np.save(out_path, np.array([int(torch.equal(got, full_ref))]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,tile_rows", [(2, 4), (2, 5), (3, 4)])
def test_gather_of_bands_gloo(tmp_path, world, tile_rows):
    out = str(tmp_path / "ok.npy")
    port = 29600 + world * 10 + tile_rows
    mp.spawn(_worker, args=(world, port, tile_rows, 8, 96, out), nprocs=world, join=True)
    assert np.load(out)[0] == 1
