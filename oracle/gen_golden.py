"""Generates tests/golden/*.npz from the REAL OpenCV routines (Python cv2) and libstdc++.

Run in the build container (needs cv2; /root/reference is read for the optional mountain
summary only).  The fixtures pin the oracle's restated OpenCV pieces and therefore the engine:
  opencv_pins.npz   gray, findHomography(4 pts), perspectiveTransform, invert, warpPerspective
  mountain.json     summary of the serial algorithm on images/mountain (seed 12345)
Usage: python oracle/gen_golden.py
"""
import json
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[0] = ROOT  # (the script dir would shadow the oracle package)
OUT = os.path.join(ROOT, "tests", "golden")


def main():
    rng = np.random.default_rng(20261018)
    d = {}
    # gray
    img = rng.integers(0, 256, (96, 128, 3), dtype=np.uint8)
    d["gray_in"] = img
    d["gray_out"] = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    # findHomography, 4 points
    srcs, dsts, Hs, oks = [], [], [], []
    for t in range(300):
        kind = t % 3
        if kind == 0:
            src = rng.integers(0, 4000, (4, 2)).astype(np.float32)
            dst = rng.integers(0, 4000, (4, 2)).astype(np.float32)
        elif kind == 1:
            src = rng.integers(0, 4000, (4, 2)).astype(np.float32)
            dst = (src + np.float32([1000, 3]) + rng.integers(-2, 3, (4, 2))).astype(np.float32)
        else:
            src = rng.integers(0, 2000, (4, 2)).astype(np.float32)
            dst = src + np.float32([777, -5])
            if t % 30 == 2:
                src[:, 0] = 5
            if t % 30 == 5:
                dst[:, 1] = 9
        H, _ = cv2.findHomography(src, dst)
        srcs.append(src); dsts.append(dst)
        oks.append(H is not None)
        Hs.append(H if H is not None else np.zeros((3, 3)))
    d["fh_src"] = np.array(srcs); d["fh_dst"] = np.array(dsts)
    d["fh_H"] = np.array(Hs); d["fh_ok"] = np.array(oks)
    # perspectiveTransform / invert
    H = np.array([[1.01, 0.02, 1900.3], [-0.01, 0.99, 4.2], [2e-6, -1e-6, 1.0]])
    pts = rng.uniform(0, 4000, (256, 2)).astype(np.float32)
    d["pt_H"] = H; d["pt_in"] = pts
    d["pt_out"] = cv2.perspectiveTransform(pts.reshape(-1, 1, 2), H).reshape(-1, 2)
    Ms = []
    for t in range(64):
        M = rng.standard_normal((3, 3)); M[2] = [rng.normal() * 1e-5, rng.normal() * 1e-5, 1]
        Ms.append(M)
    d["inv_in"] = np.array(Ms)
    d["inv_out"] = np.array([cv2.invert(M)[1] for M in Ms])
    # warpPerspective (INTER_LINEAR, BORDER_CONSTANT 0)
    src = rng.integers(0, 256, (120, 160, 3), dtype=np.uint8)
    d["warp_src"] = src
    cases = [(np.array([[1.0, 0, 50.0], [0, 1, 0], [0, 0, 1]]), (260, 120)),
             (np.array([[0.98, 0.03, 60.5], [-0.02, 1.01, 10.25], [1e-5, -2e-5, 1.0]]), (270, 150)),
             (np.array([[1.2, 0.1, -30.5], [0.05, 0.9, 20.75], [1e-4, 2e-4, 1.0]]), (173, 141)),
             (np.array([[1.0, 0, 7.5], [0, 1, 3.5], [0, 0, 1]]), (200, 13))]
    for i, (M, ds) in enumerate(cases):
        d["warp_M%d" % i] = M
        d["warp_out%d" % i] = cv2.warpPerspective(src, M, ds)
    d["opencv_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(OUT, "opencv_pins.npz"), **d)
    print("wrote opencv_pins.npz")

    # mountain summary through the oracle (decoded by cv2.imread here)
    ref = "/root/reference/images/mountain"
    if os.path.isdir(ref):
        from oracle.oracle import Oracle
        O = Oracle()
        l = cv2.imread(os.path.join(ref, "mountain1.jpg")); r = cv2.imread(os.path.join(ref, "mountain2.jpg"))
        kl, kr = O.detect(l), O.detect(r)
        m = O.match(kr, kl, r, l)
        rr = O.ransac(kr, kl, m, seed=12345)
        ok, geom, _ = O.canvas_geometry(l.shape[1], l.shape[0], r.shape[1], r.shape[0], rr["H"])
        summ = dict(kl=len(kl), kr=len(kr), m=len(m), best=rr["best_count"], best_iter=rr["best_iter"],
                    draws=rr["draws"], canvas=geom, H=rr["H"].tolist(),
                    samples_first3=rr["samples"][:3].tolist(), counts_first8=rr["counts"][:8].tolist(),
                    kl_first5=kl[:5].tolist(), seed=12345)
        json.dump(summ, open(os.path.join(OUT, "mountain.json"), "w"), indent=1)
        print("wrote mountain.json", summ["kl"], summ["kr"], summ["m"], summ["best"])


if __name__ == "__main__":
    main()
