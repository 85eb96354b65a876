"""CPU-only fuzz: the oracle restatement against the reference's own code (oracle/_ref) on random inputs.

    python tools/fuzz_oracle_vs_ref.py --cases 200 --seed 1

Random sizes (including tiny and odd ones), scene generators (textured synthetic pairs, pure noise, flat, sparse
blobs, saturated), RANSAC seeds; compares keypoints, matches (indices and SSDs), the homography bit for bit, the
status and the canvas of the whole pair.  Prints one JSON line; exit code 1 on the first difference (its inputs are
described so that the case can be replayed).  The fixed cases of tests/test_oracle_ref.py came out of runs of this."""
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


def scene(rng, synth):
    kind = rng.choice(["pair", "pair", "pair", "noise", "flat", "blobs", "saturated", "shifted"])
    w, h = int(rng.integers(12, 420)), int(rng.integers(12, 300))
    if kind == "pair":
        w, h = max(w, 96), max(h, 64)
        l, r, _ = synth.make_pair(w, h, seed=int(rng.integers(1, 1 << 30)))
    elif kind == "noise":
        l = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        r = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    elif kind == "flat":
        l = np.full((h, w, 3), int(rng.integers(0, 256)), np.uint8)
        r = l.copy()
    elif kind == "blobs":
        l = np.zeros((h, w, 3), np.uint8)
        for _ in range(int(rng.integers(1, 40))):
            x, y = int(rng.integers(0, w)), int(rng.integers(0, h))
            l[max(0, y - 3):y + 3, max(0, x - 3):x + 3] = rng.integers(0, 256, 3)
        r = np.roll(l, int(rng.integers(-20, 20)), axis=1)
    elif kind == "saturated":
        l = (rng.integers(0, 2, (h, w, 1), dtype=np.uint8) * 255).repeat(3, axis=2)
        r = np.roll(l, 3, axis=1)
    else:   # the same noise image shifted: many exact matches, SSD ties
        base = rng.integers(0, 256, (h, w + 40, 3), dtype=np.uint8)
        s = int(rng.integers(1, 40))
        l, r = base[:, :w].copy(), base[:, s:s + w].copy()
    return kind, np.ascontiguousarray(l), np.ascontiguousarray(r)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=100)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    from oracle import ref as refmod
    from oracle.oracle import Oracle
    synth = importlib.import_module(PKG + ".synth")
    O, R = Oracle(), refmod.Reference()
    rng = np.random.default_rng(a.seed)
    t0 = time.time()
    stats = {"cases": 0, "stitched": 0, "no_matches": 0, "kinds": {}}
    for case in range(a.cases):
        kind, l, r = scene(rng, synth)
        rs = int(rng.integers(1, 1 << 31))
        where = {"case": case, "kind": kind, "shape": list(l.shape), "ransac_seed": rs, "fuzz_seed": a.seed}

        def fail(what):
            print(json.dumps({"ok": False, "difference": what, **where}))
            sys.exit(1)
        kl, kr = O.detect(l), O.detect(r)
        if not (np.array_equal(kl, R.detect(l)) and np.array_equal(kr, R.detect(r))):
            fail("keypoints")
        m, mr = O.match(kr, kl, r, l), R.match(kr, kl, r, l)
        if not (np.array_equal(m["queryIdx"], mr["queryIdx"]) and np.array_equal(m["trainIdx"], mr["trainIdx"])
                and np.array_equal(m["distance"], mr["distance"])):
            fail("matches")
        if len(m) >= 1:
            Ho, Hr = O.ransac(kr, kl, m, seed=rs)["H"], R.ransac(kr, kl, mr, seed=rs)
            if (Ho is None) != (Hr is None):
                fail("ransac status")
            if Ho is not None and not np.array_equal(np.asarray(Ho, np.float64).view(np.uint64), Hr.view(np.uint64)):
                fail("homography bits")
        if rng.random() < 0.25:     # other detector / matcher options (ref: HarrisCornerOptions)
            k, th, nb, patch = (float(rng.choice([0.04, 0.06, 0.1])), float(rng.choice([1e5, 1e6, 1e7])),
                                int(rng.choice([3, 5, 7])), int(rng.choice([1, 3, 5])))
            ko, kq = O.detect(l, k=k, thresh=th, nbhd=nb), O.detect(r, k=k, thresh=th, nbhd=nb)
            if not (np.array_equal(ko, R.detect(l, k=k, thresh=th, nbhd=nb)) and np.array_equal(kq, R.detect(r, k=k, thresh=th, nbhd=nb))):
                fail("keypoints with options k=%g thresh=%g nbhd=%d" % (k, th, nb))
            mo, mq = O.match(kq, ko, r, l, patch=patch), R.match(kq, ko, r, l, patch=patch)
            if not (np.array_equal(mo["queryIdx"], mq["queryIdx"]) and np.array_equal(mo["trainIdx"], mq["trainIdx"])
                    and np.array_equal(mo["distance"], mq["distance"])):
                fail("matches with patch=%d" % patch)
            stats["option_cases"] = stats.get("option_cases", 0) + 1
        if kind == "pair" and rng.random() < 0.2:     # a three-image fold (ref: stitchAllImages)
            views = synth.make_strip(n=3, w=int(rng.integers(120, 260)), h=int(rng.integers(90, 180)), seed=int(rng.integers(1, 1 << 30)))
            po, flog = O.stitch_fold(views, seed=rs)
            pr = R.stitch_all(views, seed=rs)
            # (-1 = a garbage homography blew the canvas past this harness's buffer on either side: nothing to compare)
            oversize = pr["status"] == -1 or any(f["status"] == -1 for f in flog)
            if oversize:
                stats["folds_oversize"] = stats.get("folds_oversize", 0) + 1
            elif pr["status"] != 1 or not np.array_equal(po, pr["canvas"]):
                np.savez(os.path.join(ROOT, "gpurun_out", "fuzz_fail_fold.npz"), v0=views[0], v1=views[1], v2=views[2], seed=rs)
                fail("fold of 3 (reference status %s, shapes %s vs %s)" % (pr["status"], po.shape, None if pr["canvas"] is None else pr["canvas"].shape))
            stats["folds"] = stats.get("folds", 0) + 1
        so, sr = O.stitch_pair(l, r, seed=rs), R.stitch_pair(l, r, seed=rs)
        if (so["status"] == 1) != (sr["status"] == 1):
            fail("stitch status %s vs %s" % (so["status"], sr["status"]))
        if sr["status"] == 1:
            if not np.array_equal(so["canvas"], sr["canvas"]):
                fail("canvas")
            stats["stitched"] += 1
        if len(m) == 0:
            stats["no_matches"] += 1
        stats["cases"] += 1
        stats["kinds"][kind] = stats["kinds"].get(kind, 0) + 1
    print(json.dumps({"ok": True, "seconds": round(time.time() - t0, 1), **stats}))


if __name__ == "__main__":
    main()
