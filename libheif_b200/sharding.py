"""Multi-GPU sharding of a grid image: independent tiles are split into contiguous tile-row bands (one per rank, one
process per GPU); the only exchange is the final gather of the finished RGB row bands to rank 0.
Mirrors how ImageItem_Grid fans tiles out to threads (libheif/image-items/grid.cc:405-453), with ranks for threads."""
from typing import List, Tuple


def band_rows(tile_rows: int, world: int) -> List[int]:
    """Number of tile rows owned by each rank (contiguous bands, earlier ranks take the remainder)."""
    return [tile_rows // world + (1 if r < tile_rows % world else 0) for r in range(world)]


def my_band(tile_rows: int, tile_cols: int, world: int, rank: int) -> Tuple[int, int, List[int]]:
    """(first tile row, number of tile rows, row-major tile indices) of `rank`."""
    rows = band_rows(tile_rows, world)
    r0 = sum(rows[:rank])
    idx = [r * tile_cols + c for r in range(r0, r0 + rows[rank]) for c in range(tile_cols)]
    return r0, rows[rank], idx


def gather_bands(band, tile_rows: int, tile_h: int, world: int, rank: int, full=None):
    """Gather per-rank row bands ([band_rows*tile_h, row_bytes] uint8 tensors on the rank's device) into `full` on rank 0.
    Uses one torch.distributed.gather (NCCL over NVLink on GPUs, gloo in the CPU tests); bands of unequal height are
    padded to the tallest one for the collective."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return band
    rows = band_rows(tile_rows, world)
    mx = max(rows) * tile_h
    row_bytes = band.shape[1]
    if band.shape[0] != mx:
        pad = torch.zeros((mx, row_bytes), dtype=band.dtype, device=band.device)
        pad[:band.shape[0]] = band
    else:
        pad = band
    outs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, outs, dst=0)
    if rank != 0:
        return None
    if full is None:
        full = torch.empty((sum(rows) * tile_h, row_bytes), dtype=band.dtype, device=band.device)
    y = 0
    for r in range(world):
        h = rows[r] * tile_h
        full[y:y + h] = outs[r][:h]
        y += h
    return full
