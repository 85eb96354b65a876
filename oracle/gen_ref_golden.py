"""Generates golden vectors by RUNNING THE REFERENCE ITSELF (oracle/_ref: the unmodified
/root/reference/src/serial/main.cpp compiled against oracle/cvshim, seeded through ref_set_seed).

Run in the build container (needs /root/reference for the build and its sample images):
    python oracle/gen_ref_golden.py [--skip-photos] [--skip-4k]
Writes
    tests/golden/ref_small.npz   full outputs (keypoints, matches, H, canvases) on small seeded inputs
    tests/golden/ref_runs.json   counts, H bits and SHA-256 digests of keypoints / matches / canvases for
                                 BASELINE.json's configs (C1 mountain, C2 oilseed fold, C3 4K synthetic
                                 pair, a 1080p pair), too large to commit as pixels
Both files travel to the GPU box, where the -m gpu tests compare the CUDA engine directly with
what the reference produced.  tests/test_oracle_ref.py also holds the oracle restatement to them.
"""
import hashlib
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[0] = ROOT
OUT = os.path.join(ROOT, "tests", "golden")
PKG = "ucb-cs267-parallel-panoramic-image-stitching_b200"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def hexbits(H):
    return [format(int(v), "016x") for v in np.ascontiguousarray(H, np.float64).view(np.uint64).ravel()]


def run_pair(R, left, right, seed):
    """every stage through the reference's own functions, then its stitchTwoImages for the canvas"""
    t0 = time.time()
    kl, kr = R.detect(left), R.detect(right)
    m = R.match(kr, kl, right, left)
    H = R.ransac(kr, kl, m, seed=seed) if len(m) else None
    s = R.stitch_pair(left, right, seed=seed)
    return dict(kl=kl, kr=kr, m=m, H=H, status=s["status"], canvas=s["canvas"], secs=time.time() - t0)


def summary(r, seed):
    d = dict(seed=seed, kl=len(r["kl"]), kr=len(r["kr"]), m=len(r["m"]), status=r["status"],
             kl_sha=sha(r["kl"]), kr_sha=sha(r["kr"]),
             m_sha=sha(np.stack([r["m"]["queryIdx"], r["m"]["trainIdx"]], 1)) if len(r["m"]) else None,
             ssd_sha=sha(r["m"]["distance"]) if len(r["m"]) else None,
             H=hexbits(r["H"]) if r["H"] is not None else None,
             H_float=np.asarray(r["H"]).ravel().tolist() if r["H"] is not None else None,
             ref_seconds=round(r["secs"], 2))
    if r["canvas"] is not None:
        d["canvas"] = [int(r["canvas"].shape[1]), int(r["canvas"].shape[0])]
        d["canvas_sha"] = sha(r["canvas"])
    return d


def main():
    from oracle.ref import Reference
    synth = importlib.import_module(PKG + ".synth")
    R = Reference()
    small, runs = {}, {}

    # ---- small seeded inputs, stored in full --------------------------------------------------
    for tag, (w, h, s) in dict(a=(480, 270, 11), b=(333, 201, 5)).items():
        left, right, _ = synth.make_pair(w, h, seed=s)
        r = run_pair(R, left, right, 12345)
        assert r["status"] == 1, (tag, r["status"])
        small["%s_size" % tag] = np.array([w, h, s])
        small["%s_kl" % tag], small["%s_kr" % tag] = r["kl"], r["kr"]
        small["%s_mq" % tag], small["%s_mt" % tag] = r["m"]["queryIdx"], r["m"]["trainIdx"]
        small["%s_ssd" % tag] = r["m"]["distance"]
        small["%s_H" % tag] = r["H"]
        small["%s_canvas" % tag] = r["canvas"]
        print("small", tag, len(r["kl"]), len(r["kr"]), len(r["m"]), r["canvas"].shape)
    views = synth.make_strip(n=3, w=400, h=240, seed=9)
    f = R.stitch_all(views, seed=7)
    assert f["status"] == 1
    small["fold_size"] = np.array([3, 400, 240, 9, 7])
    small["fold_canvas"] = f["canvas"]
    # a flat image has no keypoints: the reference returns an empty Mat ("Not enough matched corners")
    flat = np.full((64, 96, 3), 77, np.uint8)
    small["flat_status"] = np.array(R.stitch_pair(flat, flat, seed=1)["status"])
    np.savez_compressed(os.path.join(OUT, "ref_small.npz"), **small)
    print("wrote ref_small.npz")

    # ---- BASELINE.json configs, digests only ---------------------------------------------------
    left, right, _ = synth.make_pair(1920, 1080, seed=31)
    runs["pair_1080p_seed31"] = summary(run_pair(R, left, right, 12345), 12345)
    print("1080p", runs["pair_1080p_seed31"])
    if "--skip-4k" not in sys.argv:
        left, right, _ = synth.make_pair(3840, 2160, seed=267)
        runs["c3_pair_4k_seed267"] = summary(run_pair(R, left, right, 12345), 12345)
        print("C3", runs["c3_pair_4k_seed267"])
    img = os.path.join(os.environ.get("PANO_REFERENCE_ROOT", "/root/reference"), "images")
    if "--skip-photos" not in sys.argv and os.path.isdir(img):
        import cv2
        l = cv2.imread(os.path.join(img, "mountain", "mountain1.jpg"))
        r = cv2.imread(os.path.join(img, "mountain", "mountain2.jpg"))
        runs["c1_mountain"] = summary(run_pair(R, l, r, 12345), 12345)
        runs["c1_mountain"]["decoder"] = "cv2.imread " + cv2.__version__
        print("C1", runs["c1_mountain"])
        ims = [cv2.imread(os.path.join(img, "oilseed", "oilseed%d.jpg" % i)) for i in (1, 2, 3, 4)]
        t0 = time.time()
        f = R.stitch_all(ims, seed=1)
        runs["c2_oilseed_fold"] = dict(seed=1, order=[1, 2, 3, 4], status=f["status"],
                                       canvas=[int(f["canvas"].shape[1]), int(f["canvas"].shape[0])],
                                       canvas_sha=sha(f["canvas"]), ref_seconds=round(time.time() - t0, 2),
                                       stage_ms=f["times_ms"], decoder="cv2.imread " + cv2.__version__)
        print("C2", runs["c2_oilseed_fold"])
    json.dump(runs, open(os.path.join(OUT, "ref_runs.json"), "w"), indent=1)
    print("wrote ref_runs.json")


if __name__ == "__main__":
    main()
