"""BASELINE config 4: synthetic n-image strip panorama in chain mode across the GPUs of one box.
Launch:  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/chain_multi_gpu.py
Each rank estimates its share of the adjacent pairs, the homographies are all-gathered (NCCL),
every rank renders its band of the canvas; rank 0 checks the result against the single-GPU
chain and prints one JSON line."""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "ucb-cs267-parallel-panoramic-image-stitching_b200"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8)
    ap.add_argument("--size", default="2000x1500")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    w, h = [int(v) for v in a.size.split("x")]
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if "PANO_NCCL_DEBUG" in os.environ:
        os.environ["NCCL_DEBUG"] = os.environ["PANO_NCCL_DEBUG"]
    else:
        os.environ.pop("NCCL_DEBUG", None)   # (any level above NONE prints the NCCL version to stdout)
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module(PKG)
    pdist = importlib.import_module(PKG + ".dist")
    synth = importlib.import_module(PKG + ".synth")
    views = synth.make_strip(n=a.n, w=w, h=h, seed=267)
    eng = pkg.Engine(device=local, seed=12345)
    times = []
    for _ in range(a.reps):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        pano, allr = pdist.stitch_chain_distributed(eng, views, device="cuda")
        dist.barrier(); torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    if rank == 0:
        t0 = time.perf_counter()
        ref, res = eng.stitchChain(views)
        t_single = time.perf_counter() - t0
        same = pano is not None and ref is not None and pano.shape == ref.shape and bool(np.array_equal(pano, ref))
        print(json.dumps({"config": "synthetic %d-image %dx%d strip, chain mode" % (a.n, w, h), "n_gpus": world,
                          "canvas": list(pano.shape[:2][::-1]) if pano is not None else None,
                          "pairs_ok": int(sum(1 for r in res if r["status"] == 0)), "identical_to_single_gpu": same,
                          "ms_distributed_best": 1000 * min(times), "ms_single_gpu": 1000 * t_single,
                          "input_MP": a.n * w * h / 1e6}), flush=True)
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
