#!/usr/bin/env python
"""copy_ceiling.py — what the box can move: the pure-copy ceiling of the end-to-end path.

The end-to-end step of bench.py moves, per 4K pair, both input images host -> device (2 x 24.9 MB) and the
canvas device -> host (~37.7 MB) through pinned memory.  This tool runs ONLY those copies — the same sizes, both
directions at once on separate streams, N GPUs concurrently (one process per GPU, like bench.py) — and prints the
aggregate rate in GB/s and in the benchmark's unit (input MP/s), so the end-to-end number can be read as a
fraction of what the box can move at that GPU count.

    python tools/copy_ceiling.py --gpus 1,2,4,8 [--numa 0|1] [--pairs 64] [--reps 5]
prints one JSON line per GPU count.
"""
import argparse
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "ucb-cs267-parallel-panoramic-image-stitching_b200"
IN_BYTES = 2 * 3840 * 2160 * 3
OUT_BYTES = 5763 * 2182 * 3


def worker(rank, world, numa, pairs, reps, barrier, out):
    import torch
    info = None
    if numa:
        info = importlib.import_module(PKG + ".numa").bind_to_gpu(rank, world)
    torch.cuda.set_device(rank)
    hin = torch.empty(IN_BYTES, dtype=torch.uint8).pin_memory()
    hout = torch.empty(OUT_BYTES, dtype=torch.uint8).pin_memory()
    hin.fill_(1)
    din = torch.empty(IN_BYTES, dtype=torch.uint8, device="cuda")
    dout = torch.zeros(OUT_BYTES, dtype=torch.uint8, device="cuda")
    up, down = torch.cuda.Stream(), torch.cuda.Stream()
    best = None
    for r in range(reps + 1):
        torch.cuda.synchronize()
        barrier.wait()
        t0 = time.perf_counter()
        for _ in range(pairs):
            with torch.cuda.stream(up):
                din.copy_(hin, non_blocking=True)
            with torch.cuda.stream(down):
                hout.copy_(dout, non_blocking=True)
        torch.cuda.synchronize()
        barrier.wait()
        dt = time.perf_counter() - t0
        if r > 0:
            best = dt if best is None else min(best, dt)
    out.put((rank, best, info))


def main():
    import torch.multiprocessing as mp
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1")
    ap.add_argument("--numa", type=int, default=1)
    ap.add_argument("--pairs", type=int, default=64)
    ap.add_argument("--reps", type=int, default=4)
    a = ap.parse_args()
    mp.set_start_method("spawn", force=True)
    for n in [int(v) for v in a.gpus.split(",")]:
        barrier, out = mp.Barrier(n), mp.Queue()
        ps = [mp.Process(target=worker, args=(r, n, a.numa, a.pairs, a.reps, barrier, out)) for r in range(n)]
        for p in ps:
            p.start()
        res = [out.get() for _ in range(n)]
        for p in ps:
            p.join()
        t = max(r[1] for r in res)          # all ranks start together: the slowest one ends the step
        gb = n * a.pairs * (IN_BYTES + OUT_BYTES) / 1e9
        print(json.dumps({"tool": "copy_ceiling", "n_gpus": n, "numa_bind": bool(a.numa), "pairs_per_gpu": a.pairs,
                          "seconds": t, "aggregate_GBps_both_directions": gb / t,
                          "h2d_GBps_per_gpu": a.pairs * IN_BYTES / 1e9 / t, "d2h_GBps_per_gpu": a.pairs * OUT_BYTES / 1e9 / t,
                          "ceiling_input_MP_per_s": n * a.pairs * 2 * 3840 * 2160 / 1e6 / t,
                          "ms_per_pair_per_gpu": 1000 * t / a.pairs,
                          "placement": [r[2] for r in sorted(res)]}), flush=True)


if __name__ == "__main__":
    main()
