"""CPU-only fuzz: the oracle's restated OpenCV pieces against the real OpenCV (cv2) on random inputs.

    python tools/fuzz_oracle_vs_cv2.py --cases 2000 --seed 1

gray (cvtColor BGR2GRAY), the 4-point findHomography (bit for bit, including degenerate quads), perspectiveTransform,
3x3 invert and warpPerspective (INTER_LINEAR, BORDER_CONSTANT) followed by the reference's overlay.  One JSON line;
exit code 1 on the first difference."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def random_h(rng, w, h):
    a = rng.uniform(-0.3, 0.3)
    s = rng.uniform(0.7, 1.4)
    H = np.array([[s * np.cos(a), -s * np.sin(a), rng.uniform(-0.6, 0.6) * w],
                  [s * np.sin(a), s * np.cos(a), rng.uniform(-0.6, 0.6) * h],
                  [rng.uniform(-1e-3, 1e-3), rng.uniform(-1e-3, 1e-3), 1.0]])
    if rng.random() < 0.2:
        H[2, 2] = rng.uniform(0.5, 2.0)
    return H


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=500)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    import cv2
    from oracle.oracle import Oracle
    O = Oracle()
    rng = np.random.default_rng(a.seed)
    t0 = time.time()
    n = {"gray": 0, "homography": 0, "homography_empty": 0, "points": 0, "invert": 0, "warp": 0, "warp_px": 0}

    def fail(what, **kw):
        print(json.dumps({"ok": False, "difference": what, "fuzz_seed": a.seed, **kw}, default=str))
        sys.exit(1)
    for case in range(a.cases):
        w, h = int(rng.integers(3, 260)), int(rng.integers(3, 200))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if rng.random() < 0.3:
            img[rng.integers(0, h):, :] = 0                      # black regions matter to the overlay
        if not np.array_equal(O.gray(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)):
            fail("gray", case=case)
        n["gray"] += 1
        for _ in range(10):                                       # 4-point findHomography
            kind = rng.random()
            src = rng.integers(0, 4000, (4, 2)).astype(np.float32)
            dst = (src + rng.integers(-60, 60, (4, 2))).astype(np.float32)
            if kind < 0.1:
                src[:, 0] = src[0, 0]                             # all on one vertical line
            elif kind < 0.2:
                dst[1] = dst[0]                                   # a repeated point
            elif kind < 0.3:
                src[2] = (src[0] + src[1]) / 2                    # three collinear points
            Hc, _ = cv2.findHomography(src, dst)
            Ho = O.find_homography4(src, dst)
            if (Hc is None) != (Ho is None):
                fail("findHomography emptiness", case=case, src=src.tolist(), dst=dst.tolist())
            if Hc is None:
                n["homography_empty"] += 1
            elif not np.array_equal(bits(Hc), bits(Ho)):
                fail("findHomography bits", case=case, src=src.tolist(), dst=dst.tolist())
            n["homography"] += 1
        H = random_h(rng, w, h)
        pts = rng.uniform(-50, 4000, (16, 2)).astype(np.float32)
        if not np.array_equal(cv2.perspectiveTransform(pts[None], H)[0].view(np.uint32), O.perspective_transform(pts, H).view(np.uint32)):
            fail("perspectiveTransform", case=case, H=H.tolist())
        n["points"] += 1
        Hi = O.invert33(H)
        ci = cv2.invert(H)[1]
        if Hi is not None and not np.array_equal(bits(ci), bits(Hi)):
            fail("invert", case=case, H=H.tolist())
        n["invert"] += 1
        cw, ch = int(rng.integers(3, 400)), int(rng.integers(3, 300))
        wc = cv2.warpPerspective(img, H, (cw, ch))
        wo = O.warp_perspective(img, H, (cw, ch))
        if not np.array_equal(wc, wo):
            d = np.argwhere((wc != wo).any(axis=2))
            fail("warpPerspective", case=case, H=H.tolist(), src=[w, h], dst=[cw, ch], n_px=len(d), first=d[0].tolist(),
                 cv=wc[tuple(d[0])].tolist(), oracle=wo[tuple(d[0])].tolist())
        n["warp"] += 1
        n["warp_px"] += cw * ch
    print(json.dumps({"ok": True, "seconds": round(time.time() - t0, 1), "cases": a.cases, **n}))


if __name__ == "__main__":
    main()
