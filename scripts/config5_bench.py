"""BASELINE config 5: a 32768 x 32768 tiled 12-bit picture (32 x 32 tiles of 1024 x 1024) of which only 64 randomly chosen tiles are
decoded (region of interest), round-robin over the GPUs.  Tiles are independent: rank r decodes picks[r::world] as one batch through
the fused C-ABI entry point (host access units -> RRGGBB_LE in page-locked host memory); no data-path collective.

    python scripts/config5_bench.py [--tile 1024 --ntiles 64 --grid 32 --steps 5 --warmup 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 scripts/config5_bench.py

Prints one JSON line (rank 0): ms per step (max over ranks, barrier + synchronize on both sides), MP/s, tiles/s, and the md5 over
the RGB of every tile in pick order -- the same for every GPU count (tests/test_multi_gpu.py compares 1 and 2 GPUs and checks
tiles against the oracle)."""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def picks_lcg(n, total, seed=0xB2005):
    s, out = seed, []
    while len(out) < n:
        s = (s * 1664525 + 1013904223) & 0xffffffff
        t = (s >> 8) % total
        if t not in out:
            out.append(t)
    return out


def make_tile(t, tile, bd):
    from libheif_b200 import hevc_enc
    y, cb, cr = hevc_enc.synthetic_image(0xB200 + t, tile, tile, bd, True)
    return hevc_enc.encode_intra(y, cb, cr, bit_depth=bd, log2_ctb_size=5, qp=27, wpp=1, seed=0xB200 + t, vui_present=1, colour_description_present=1,
                                 colour_primaries=9, transfer_characteristics=16, matrix_coefficients=9, full_range=0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tile", type=int, default=1024)
    ap.add_argument("--ntiles", type=int, default=64)
    ap.add_argument("--grid", type=int, default=32)
    ap.add_argument("--bit-depth", type=int, default=12)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--dump-dir", default="", help="write every tile's RGB as <dir>/tile_<index>.rgb (tests)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    import torch
    import torch.distributed as dist
    import libheif_b200 as lb
    from concurrent.futures import ThreadPoolExecutor
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    picks = picks_lcg(args.ntiles, args.grid * args.grid)
    mine = picks[rank::world]
    with ThreadPoolExecutor(max(1, min(16, (os.cpu_count() or 16) // world))) as ex:
        tiles = list(ex.map(lambda t: make_tile(t, args.tile, args.bit_depth), mine))
    T = args.tile
    dec = lb.Decoder(host_threads=max(1, 16 // world))
    out = torch.empty((T, max(1, len(mine)) * T * 6), dtype=torch.uint8, pin_memory=True)

    def step():
        if mine:
            dec.decode_grid_to_rgb_host(tiles, len(mine), 1, lb.CHROMA_INTERLEAVED_RRGGBB_LE, out=out.numpy())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    barrier()
    ms = (time.perf_counter() - t0) * 1e3 / args.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    st = dec.stats() if mine else None
    o = out.numpy()
    md5s = {}
    for k, t in enumerate(mine):
        rgb = np.ascontiguousarray(o[:, k * T * 6:(k + 1) * T * 6])
        md5s[t] = hashlib.md5(rgb.tobytes()).hexdigest()
        if args.dump_dir:
            rgb.tofile(os.path.join(args.dump_dir, f"tile_{t}.rgb"))
    if world > 1:
        allm = [None] * world
        dist.all_gather_object(allm, md5s)
        md5s = {k: v for m in allm for k, v in m.items()}
    if rank == 0:
        h = hashlib.md5("".join(md5s[t] for t in picks).encode()).hexdigest()
        px = args.ntiles * T * T
        print(json.dumps({"workload": f"config 5: {args.ntiles} LCG-picked {T}x{T} {args.bit_depth}-bit tiles of a {args.grid}x{args.grid} tile grid -> RRGGBB_LE, round-robin over the GPUs",
                          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "mp_s": px / 1e6 / (ms / 1e3), "tiles_per_s": args.ntiles / (ms / 1e3),
                          "rank0": {"tiles": len(mine), "entropy_ms": st.entropy_ms if st else None, "recon_ms": st.recon_ms if st else None, "front_end": st.front_end if st else None},
                          "api": "b200_decode_grid_to_rgb_host per rank (host access units -> page-locked host RRGGBB_LE)", "md5_of_tile_md5s": h}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
