#!/usr/bin/env python
"""bench.py -- megapixels/s of HEIC-grid decode -> RGB on B200 (BASELINE.json metric), one JSON line on stdout.

Workload (BASELINE configs[2]): a synthetic 16384x16384 HEIC grid = 256 independent 1024x1024 HEVC-intra tiles
(8-bit 4:2:0, fixed QP 27, CTB 32, WPP, SAO + deblocking on; seed 0xB200 + tile index) decoded to
interleaved RGB24.  One "step" = the whole grid once.
  value : device-timed (CUDA events) MP/s with the command stream (post-CABAC) resident in HBM: reconstruction +
          deblocking + SAO/paste + colour conversion (+ the NCCL gather of the RGB bands when N > 1).
  e2e   : host bitstreams -> host RGB through the C ABI: CABAC parse on the host cores, H2D of the command stream,
          kernels, gather, D2H of the RGB into pinned host memory, all inside the timed region.
  --impl reference : the reference CPU path on this box's host cores (FFmpeg HEVC decode in the libde265 role per tile,
          tiles over all cores like ImageItem_Grid does, paste, then the UNMODIFIED reference convert_colorspace from
          oracle/_ref/libheif_ref.so), on a bounded sample of the same tiles.
Multi-GPU (torchrun, one rank per GPU): tile rows are sharded across ranks (strong scaling: the grid is fixed), the
only collective is the final gather of RGB row bands to rank 0 (NCCL over NVLink).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TILE = 1024
QP = 27


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_tile(idx, tile=TILE):
    from libheif_b200 import hevc_enc
    y, cb, cr = hevc_enc.synthetic_image(0xB200 + idx, tile, tile, 8, True)
    return hevc_enc.encode_intra(y, cb, cr, bit_depth=8, log2_ctb_size=5, qp=QP, wpp=1, seed=0xB200 + idx, vui_present=1,
                                 colour_description_present=1, colour_primaries=1, transfer_characteristics=13,
                                 matrix_coefficients=6, full_range=0)


def make_tiles(indices, tile=TILE, workers=None):
    workers = workers or min(64, effective_cores())
    with ThreadPoolExecutor(workers) as ex:
        return list(ex.map(lambda i: make_tile(i, tile), indices))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.samples = []
        self.stop = False
        self.idx = gpu_index
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                r = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.idx)],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5)
                f = [x.strip() for x in r.stdout.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(float(s[0])) for s in self.samples)
        reasons = set()
        for s in self.samples:
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(self.samples[0][1])), "reasons": sorted(reasons), "samples": len(sm)}


def effective_cores():
    """Host cores this process may actually use: min(visible CPUs, cgroup CPU quota)."""
    n = os.cpu_count() or 1
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(int(q) / int(p))))
    except Exception:
        pass
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    return n


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


# ------------------------------------------------------------------------------------------ reference CPU arm
def reference_arm(args, tiles_side, sample_side, steps, warmup, cores):
    """FFmpeg tile decode on all host cores + paste + unmodified reference colour conversion; returns MP/s per step."""
    from oracle import bindings as ob
    if ob.ref_plugin() is None or ob.avcodec_dir() is None:
        return None, "oracle/_ref reference build or FFmpeg missing"
    idx = [r * tiles_side + c for r in range(sample_side) for c in range(sample_side)]
    tiles = make_tiles(idx)
    ob.ffmpeg_decode(tiles[0])        # loads libavcodec
    W = H = sample_side * TILE

    def step():
        with ThreadPoolExecutor(cores) as ex:
            dec = list(ex.map(lambda t: ob.ffmpeg_decode(t, 1)[0], tiles))
        y = np.empty((H, W), np.uint16); cb = np.empty((H // 2, W // 2), np.uint16); cr = np.empty((H // 2, W // 2), np.uint16)
        for k, pl in enumerate(dec):              # copy_image_to (grid.cc:574)
            c, r = k % sample_side, k // sample_side
            y[r * TILE:(r + 1) * TILE, c * TILE:(c + 1) * TILE] = pl[0]
            cb[r * TILE // 2:(r + 1) * TILE // 2, c * TILE // 2:(c + 1) * TILE // 2] = pl[1]
            cr[r * TILE // 2:(r + 1) * TILE // 2, c * TILE // 2:(c + 1) * TILE // 2] = pl[2]
        out, ow, oh, _ = ob.ref_postprocess(y, cb, cr, None, 1, 8, (1, 13, 6, 0), [], 10)
        return out

    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
    ms = 1e3 * sum(ts) / len(ts)
    return {"mp_s": W * H / 1e6 / (ms / 1e3), "ms": ms, "sample": f"{sample_side}x{sample_side} tiles of the same grid ({W}x{H}, {W * H / 1e6:.1f} MP) per step",
            "pixels": W * H}, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tiles-side", type=int, default=16, help="grid is tiles-side x tiles-side tiles of 1024x1024 (16 = BASELINE config)")
    ap.add_argument("--ref-sample-side", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--front-end", default="device", choices=["device", "host"], help="where CABAC runs: GPU (one warp per WPP sub-stream) or host cores")
    args = ap.parse_args()
    warmup = max(3, args.warmup)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    side = args.tiles_side
    cores = effective_cores()
    workload = f"{side * TILE}x{side * TILE} HEIC grid, {side * side} x {TILE}x{TILE} HEVC-intra tiles, 8-bit 4:2:0 -> RGB24, QP {QP}, CTB 32, WPP, SAO+deblock"

    if args.impl == "reference":
        if rank != 0:
            return
        res, why = reference_arm(args, side, min(side, args.ref_sample_side), args.steps, min(warmup, 1), cores)
        if res is None:
            print(json.dumps({"impl": "reference", "unavailable": why}))
            return
        line = {"impl": "reference", "metric": "megapixels/sec HEIC-grid decode->RGB", "value": res["mp_s"], "unit": "MP/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": min(warmup, 1), "ms_per_step": res["ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic", "config": {"workload": workload, "sample": res["sample"]},
                "cpu_baseline": {"value": res["mp_s"], "unit": "MP/s", "cores": cores, "kind": "reference", "sample": res["sample"]},
                "e2e": {"value": res["mp_s"], "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import libheif_b200 as lb
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # ---- shard: contiguous tile-row bands per rank (libheif_b200/sharding.py, gloo-tested in tests/test_sharding.py)
    from libheif_b200 import sharding
    r0, nrows, my_idx = sharding.my_band(side, side, world, rank)
    t_gen = time.perf_counter()
    tiles = make_tiles(my_idx, workers=max(4, cores // world))
    t_gen = time.perf_counter() - t_gen
    W, H = side * TILE, side * TILE
    band_h = nrows * TILE
    dec = lb.Decoder(host_threads=max(1, cores // world))
    dec.set_front_end(args.front_end == "device")
    band = torch.empty((max(band_h, 1), W * 3), dtype=torch.uint8, device=dev)
    full = torch.empty((H, W * 3), dtype=torch.uint8, device=dev) if (rank == 0 and world > 1) else None
    host_out = torch.empty((H, W * 3), dtype=torch.uint8, pin_memory=True) if rank == 0 else None
    stream = torch.cuda.current_stream()

    def gather():
        return sharding.gather_bands(band, side, TILE, world, rank, full) if world > 1 else band

    def device_step():
        if nrows:
            dec.rerun_device(stream)
            dec.to_rgb_device(lb.CHROMA_INTERLEAVED_RGB, out=band, stream=stream)
        return gather()

    def e2e_step():
        if nrows:
            dec.decode_grid(tiles, cols=side, rows=nrows, stream=stream)
            dec.to_rgb_device(lb.CHROMA_INTERLEAVED_RGB, out=band, stream=stream)
        res = gather()
        if rank == 0:
            host_out.copy_(res, non_blocking=True)
        torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up (also uploads the command stream for the device leg)
    for _ in range(warmup):
        e2e_step()
    st0 = dec.stats() if nrows else None
    with ClockSampler(local_rank) as clk:
        # ---- leg A: kernels with the command stream resident in HBM (CUDA events, max over ranks)
        for _ in range(warmup):
            device_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            device_step()
        e1.record(stream)
        barrier()
        dev_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        # ---- leg B: end to end through the C ABI, host buffers in, pinned host RGB out (wall clock around synchronised region)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    # ---- per-kernel device times (outside the timed regions): average over a few launches, CUDA events
    kern = {"entropy": 0.0, "recon": 0.0, "deblock": 0.0, "sao_paste": 0.0, "k6_colour": 0.0}
    nk = 5
    overlapped = False
    if nrows:
        for _ in range(nk):
            dec.rerun_device(stream)
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record(stream); dec.to_rgb_device(lb.CHROMA_INTERLEAVED_RGB, out=band, stream=stream); k1.record(stream)
            torch.cuda.synchronize()
            s = dec.stats()
            overlapped = s.front_end == 2
            kern["entropy"] += s.entropy_ms / nk; kern["recon"] += s.recon_ms / nk; kern["deblock"] += s.deblock_ms / nk; kern["sao_paste"] += s.sao_ms / nk; kern["k6_colour"] += k0.elapsed_time(k1) / nk
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- rank 0: report
    pixels = W * H
    my_px = band_h * W
    peak, peak_src = measured_peak()
    C = st0.command_bytes / my_px                                     # measured command-stream bytes per pixel
    bpp = st0.bitstream_bytes / my_px
    # algorithmic B/px: entropy reads the bitstream and writes the command stream once
    alg = {"entropy": bpp + C, "recon": 1.5 + C, "deblock": 3.0, "sao_paste": 3.0, "k6_colour": 4.5}
    if args.front_end == "host":
        kern.pop("entropy")
    if overlapped:                                   # K0 and K1 ran concurrently: one time for both, one algorithmic figure
        kern["entropy+recon"] = kern.pop("entropy") + kern.pop("recon")
        alg["entropy+recon"] = bpp + 1.5 + C         # bitstream read, planes written; the command stream stays in L2 / HBM in between
    dom = max(kern, key=kern.get)
    ach = alg[dom] * my_px / (kern[dom] * 1e-3) / 1e9
    # dram__bytes of the dominant kernel from the committed ncu capture of this exact workload (profiles/r01_k0_entropy_ncu_256tiles_v3.txt:
    # 5.52 GB/s over 96.9 ms); other workloads / kernels: not captured
    traffic = 5.35e8 if (dom == "entropy" and world == 1 and not overlapped) else None
    stats_e2e = dec.stats()
    line = {
        "metric": "megapixels/sec HEIC-grid decode->RGB", "value": pixels / 1e6 / (dev_ms / 1e3), "unit": "MP/s", "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload, "sharding": f"{world} x contiguous tile-row bands, NCCL gather of RGB bands to rank 0" if world > 1 else "single GPU",
                   "l2": "inputs larger than L2 (command stream + planes > 126 MB per GPU)" if st0.command_bytes + my_px * 1.5 > 126e6 else "flush not needed: see note",
                   "bits_per_pixel": 8.0 * st0.bitstream_bytes / my_px, "command_bytes_per_pixel": C, "host_parser_threads": max(1, cores // world), "front_end": args.front_end + (" (CABAC on the GPU, one warp per WPP sub-stream)" if args.front_end == "device" else " (CABAC on the host cores)")},
        "e2e": {"value": pixels / 1e6 / (e2e_ms / 1e3), "unit": "MP/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(stats_e2e.h2d_bytes * (pixels / my_px)),
                "d2h_bytes_per_step": pixels * 3, "host_parse_ms": stats_e2e.parse_ms, "host_pack_ms": stats_e2e.pack_ms},
        "gpu_launches": ((6 if args.front_end == "device" else 4) + 1) * args.steps,
        "clocks": clk.summary(),
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes_per_pixel": alg[dom],
                     "kernels_ms": kern, "kernels_gb_s": {k: alg[k] * my_px / (v * 1e-3) / 1e9 if v > 0 else None for k, v in kern.items()},
                     "pipeline_A_bytes_per_pixel": 12.0 + C,
                     "pipeline_gb_s": (12.0 + C) * my_px / (sum(kern.values()) * 1e-3) / 1e9 if sum(kern.values()) > 0 else None},
        "setup": {"tile_generation_s": t_gen},
    }
    if world == 1 and not args.no_cpu_baseline:
        res, why = reference_arm(args, side, min(side, args.ref_sample_side), 2, 1, cores)
        line["cpu_baseline"] = ({"value": res["mp_s"], "unit": "MP/s", "cores": cores, "kind": "reference", "sample": res["sample"]}
                                if res else {"value": None, "unit": "MP/s", "cores": cores, "kind": "reference", "sample": f"unavailable: {why}"})
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
