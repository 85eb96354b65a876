"""Runs the fused pair pipeline on one synthetic pair a few times and prints the per-stage
device times (CUDA events inside the C ABI).  Used plain for a breakdown and under ncu for the
per-kernel launch list / full captures (see profiles/)."""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "ucb-cs267-parallel-panoramic-image-stitching_b200"


def cached_pair(w, h, seed):
    synth = importlib.import_module(PKG + ".synth")
    d = os.environ.get("PANO_SYNTH_CACHE", "/tmp/pano_synth_cache")
    os.makedirs(d, exist_ok=True)
    f = os.path.join(d, "pair_%dx%d_%d.npz" % (w, h, seed))
    if os.path.exists(f):
        z = np.load(f)
        return z["left"], z["right"], z["H"]
    left, right, H = synth.make_pair(w, h, seed=seed)
    np.savez(f, left=left, right=right, H=H)
    return left, right, H


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="3840x2160")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--seed", type=int, default=267)
    ap.add_argument("--matcher", type=int, default=0)
    ap.add_argument("--replay", type=int, default=0, help="0 chunked, 1 resident")
    a = ap.parse_args()
    w, h = [int(v) for v in a.size.split("x")]
    import torch
    pkg = importlib.import_module(PKG)
    left, right, _ = cached_pair(w, h, a.seed)
    L, R = torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()
    eng = pkg.Engine(0, 12345)
    eng.set_matcher(a.matcher)
    eng.set_replay_mode(a.replay)
    for i in range(a.reps):
        n0 = eng.kernel_launches()
        _, r = eng.stitchTwoImages(L, R, fetch=False)
        r["H"] = r["H"].tolist()
        r["launches"] = eng.kernel_launches() - n0
        print(json.dumps(r))
    eng.close()


if __name__ == "__main__":
    main()
