#!/usr/bin/env python3
"""Quality scoring of a stitched panorama against a ground-truth panorama.

Same metrics, argument order (baseline, test) and poor / acceptable / good bands as the
reference's evaluator (ref: evaluate_panorama.py:6-119), re-implemented without scikit-image
(not installed here): the test image is registered to the baseline with ORB + Hamming brute
force + RANSAC homography, then PSNR, SSIM, inlier ratio, mean reprojection error and seam
smoothness are computed.  SSIM follows scikit-image's defaults (7x7 uniform window, sample
covariance, K1 = 0.01, K2 = 0.03, data range 255, mean over channels and pixels); the reference
passes a `mask=` keyword that scikit-image does not act on, so the whole-frame value is the one
reported as "SSIM" and the overlap-only value is printed next to it.
"""
import argparse
import sys

import cv2
import numpy as np

BANDS = {  # metric: (acceptable, good, higher_is_better)
    "PSNR": (25.0, 35.0, True),
    "SSIM": (0.80, 0.90, True),
    "Inlier Ratio": (0.50, 0.70, True),
    "Reprojection Error": (3.0, 1.0, False),
    "Seam Smoothness": (30.0, 10.0, False),
}


def ssim_map(a, b, win=7, data_range=255.0):
    a = a.astype(np.float64); b = b.astype(np.float64)
    n = win * win
    cov_norm = n / (n - 1.0)
    k = (win, win)
    box = lambda x: cv2.blur(x, k, borderType=cv2.BORDER_REFLECT)
    ua, ub = box(a), box(b)
    uaa, ubb, uab = box(a * a), box(b * b), box(a * b)
    va, vb, vab = cov_norm * (uaa - ua * ua), cov_norm * (ubb - ub * ub), cov_norm * (uab - ua * ub)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    s = ((2 * ua * ub + c1) * (2 * vab + c2)) / ((ua * ua + ub * ub + c1) * (va + vb + c2))
    pad = (win - 1) // 2
    return s, pad


def register(base, test, thr):
    orb = cv2.ORB_create(5000)
    k1, d1 = orb.detectAndCompute(base, None)
    k2, d2 = orb.detectAndCompute(test, None)
    if d1 is None or d2 is None:
        raise RuntimeError("no ORB features")
    m = cv2.BFMatcher(cv2.NORM_HAMMING).match(d1, d2)
    if len(m) < 4:
        raise RuntimeError("Not enough matches for homography")
    p1 = np.float32([k1[x.queryIdx].pt for x in m]); p2 = np.float32([k2[x.trainIdx].pt for x in m])
    H, mask = cv2.findHomography(p1, p2, cv2.RANSAC, thr)
    if H is None:
        raise RuntimeError("Homography estimation failed")
    mask = mask.ravel().astype(bool)
    proj = cv2.perspectiveTransform(p1.reshape(-1, 1, 2).astype(np.float64), H).reshape(-1, 2)
    err = float(np.linalg.norm(proj[mask] - p2[mask], axis=1).mean())
    return H, float(mask.mean()), err


def compute_metrics(base, test, thr=3.0):
    H, inlier_ratio, reproj = register(base, test, thr)
    h, w = test.shape[:2]
    warped = cv2.warpPerspective(base, H, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
    overlap = np.any(warped != 0, axis=2)
    if not overlap.any():
        raise RuntimeError("No overlap region found")
    d = warped.astype(np.float32) - test.astype(np.float32)
    mse = float(np.mean(d[overlap] ** 2))
    psnr = 10 * np.log10(255.0 ** 2 / mse) if mse > 0 else float("inf")
    per_ch, per_ch_ov = [], []
    for c in range(3):
        s, pad = ssim_map(warped[:, :, c], test[:, :, c])
        inner = s[pad:h - pad, pad:w - pad]
        per_ch.append(inner.mean())
        per_ch_ov.append(inner[overlap[pad:h - pad, pad:w - pad]].mean())
    g = cv2.cvtColor(cv2.absdiff(warped, test), cv2.COLOR_BGR2GRAY)
    mag = np.hypot(cv2.Sobel(g, cv2.CV_64F, 1, 0), cv2.Sobel(g, cv2.CV_64F, 0, 1))
    ring = cv2.dilate(overlap.astype(np.uint8), np.ones((3, 3), np.uint8)).astype(bool) & ~overlap
    seam = float(mag[ring].mean()) if ring.any() else 0.0
    return {"PSNR": float(psnr), "SSIM": float(np.mean(per_ch)), "Inlier Ratio": inlier_ratio,
            "Reprojection Error": reproj, "Seam Smoothness": seam}, float(np.mean(per_ch_ov))


def band(name, v):
    acc, good, hib = BANDS[name]
    ok_good = v >= good if hib else v <= good
    ok_acc = v >= acc if hib else v <= acc
    return "good" if ok_good else "acceptable" if ok_acc else "poor"


def report(metrics, ssim_overlap=None, out=sys.stdout):
    print("Quality Levels: good / acceptable / poor (thresholds per metric)", file=out)
    for name, (acc, good, hib) in BANDS.items():
        op = ">=" if hib else "<="
        print("  %-19s good %s %g, acceptable %s %g" % (name, op, good, op, acc), file=out)
    print(file=out)
    cats = []
    for name, v in metrics.items():
        c = band(name, v)
        cats.append(c)
        print("%-19s: %.4f [%s]" % (name, v, c), file=out)
    if ssim_overlap is not None:
        print("%-19s: %.4f (overlap region only; informational)" % ("SSIM (masked)", ssim_overlap), file=out)
    overall = "Poor" if "poor" in cats else "Acceptable" if "acceptable" in cats else "Good"
    print("\nOverall stitching quality: %s" % overall, file=out)
    return overall


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("baseline", help="Baseline panorama image")
    ap.add_argument("test", help="Test panorama image")
    ap.add_argument("--threshold", type=float, default=3.0, help="RANSAC reproj threshold in pixels")
    a = ap.parse_args()
    base, test = cv2.imread(a.baseline), cv2.imread(a.test)
    if base is None or test is None:
        raise RuntimeError("Failed to load images")
    m, so = compute_metrics(base, test, a.threshold)
    report(m, so)


if __name__ == "__main__":
    main()
