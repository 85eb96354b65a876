"""CPU-only fuzz: the engine's device-side arithmetic (csrc/pano_core.cuh, the functions the CUDA kernels call,
compiled for the host by tests/hostsim) against the oracle on random inputs.

    python tools/fuzz_device_core_vs_oracle.py --cases 2000 --seed 1

find_homography4 (the dlt_kernel's solver incl. OpenCV's Jacobi) bit for bit, canvas_geometry, the fixed-point warp
model (warp_coord + warp_pixel), and both checked-Newton coordinate paths of the warp kernels against the exact IEEE
expression (no accepted fast result may differ).  One JSON line; exit code 1 on the first difference."""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "ucb-cs267-parallel-panoramic-image-stitching_b200"


def load_hostsim():
    d = os.path.join(ROOT, "tests", "hostsim")
    so, src = os.path.join(d, "libhostsim.so"), os.path.join(d, "hostsim.cpp")
    hdrs = [os.path.join(ROOT, PKG, "csrc", f) for f in ("pano_core.cuh", "replay_plan.hpp")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src])
    return C.CDLL(so)


def p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def random_h(rng, w, h):
    a, s = rng.uniform(-0.3, 0.3), rng.uniform(0.7, 1.4)
    return np.array([[s * np.cos(a), -s * np.sin(a), rng.uniform(-0.6, 0.6) * w],
                     [s * np.sin(a), s * np.cos(a), rng.uniform(-0.6, 0.6) * h],
                     [rng.uniform(-1e-3, 1e-3), rng.uniform(-1e-3, 1e-3), rng.uniform(0.5, 2.0) if rng.random() < 0.2 else 1.0]])


def fast_path_ok(M, cw, ch):
    """the domain warp.cu::fast_path_ok admits the fast coordinate path for (restated): positive denominators on the
    canvas corners, OpenCV's 64-px coordinate blocks, |coordinate| * 32 below the magic-rounding range"""
    if cw < 64 or ch < 16:
        return False
    wmin, nmax = 1e300, 0.0
    for cx, cy in ((0.0, 0.0), (cw, 0.0), (cw, ch), (0.0, ch)):
        wmin = min(wmin, M[6] * cx + M[7] * cy + M[8])
        nmax = max(nmax, abs(M[0] * cx + M[1] * cy + M[2]), abs(M[3] * cx + M[4] * cy + M[5]))
    return wmin > 1e-9 and 32.0 * nmax / wmin < 4000000.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=500)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    from oracle.oracle import Oracle
    O, hs = Oracle(), load_hostsim()
    rng = np.random.default_rng(a.seed)
    t0 = time.time()
    n = {"homography": 0, "homography_empty": 0, "geometry": 0, "warp": 0, "warp_px": 0, "fast_px": 0, "fast_exact_path_px": 0}

    def fail(what, **kw):
        print(json.dumps({"ok": False, "difference": what, "fuzz_seed": a.seed, **kw}, default=str))
        sys.exit(1)
    for case in range(a.cases):
        for _ in range(20):
            kind = rng.random()
            src = rng.integers(0, 4000, (4, 2)).astype(np.float32)
            dst = (src + rng.integers(-60, 60, (4, 2))).astype(np.float32)
            if kind < 0.1:
                src[:, 1] = src[0, 1]
            elif kind < 0.2:
                dst[3] = dst[2]
            elif kind < 0.3:
                src[2] = (src[0] + src[1]) / 2
            H = np.empty((3, 3), np.float64)
            ok = hs.hs_find_homography4(p(src, C.c_float), p(dst, C.c_float), p(H, C.c_double))
            Ho = O.find_homography4(src, dst)
            if bool(ok) != (Ho is not None):
                fail("find_homography4 emptiness", case=case, src=src.tolist(), dst=dst.tolist())
            if Ho is None:
                n["homography_empty"] += 1
            elif not np.array_equal(bits(H), bits(Ho)):
                fail("find_homography4 bits", case=case, src=src.tolist(), dst=dst.tolist())
            n["homography"] += 1
        wl, hl, wr, hr = (int(v) for v in rng.integers(16, 5000, 4))
        H = random_h(rng, wr, hr)
        geom, TH, Minv = np.zeros(6, np.int32), np.empty(9), np.empty(9)
        hs.hs_canvas_geometry(wl, hl, wr, hr, p(H, C.c_double), p(geom, C.c_int32), p(TH, C.c_double), p(Minv, C.c_double))
        ok, g, THo = O.canvas_geometry(wl, hl, wr, hr, H)
        if bool(geom[5]) != ok or (ok and (tuple(int(v) for v in geom[:4]) != g or not np.array_equal(bits(TH), bits(THo).ravel()))):
            fail("canvas_geometry", case=case, H=H.tolist(), sizes=[wl, hl, wr, hr])
        n["geometry"] += 1
        if ok and 0 < geom[0] < 30000 and 0 < geom[1] < 30000 and fast_path_ok(Minv, int(geom[0]), int(geom[1])):
            # both Newton paths on this canvas (inside the domain the host admits them for), a sparse row sample
            for variant in (1, 2):
                out = np.zeros(3, np.uint64)
                hs.hs_warp_fast_check(p(np.ascontiguousarray(Minv), C.c_double), int(geom[0]), int(geom[1]),
                                      max(1, int(geom[1]) // 24), 3, p(out, C.c_uint64), variant)
                if out[0] != 0:
                    fail("warp_coord_fast%s accepted a wrong coordinate" % ("2" if variant == 2 else ""), case=case, H=H.tolist())
                n["fast_px"] += int(out[2])
                n["fast_exact_path_px"] += int(out[1])
        w, h = int(rng.integers(3, 200)), int(rng.integers(3, 160))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        M = random_h(rng, w, h)
        cw, ch = int(rng.integers(3, 300)), int(rng.integers(3, 220))
        dst = np.empty((ch, cw, 3), np.uint8)
        hs.hs_warp_perspective(p(img, C.c_uint8), w, h, C.c_size_t(img.strides[0]), p(np.ascontiguousarray(M), C.c_double),
                               p(dst, C.c_uint8), cw, ch, C.c_size_t(dst.strides[0]))
        if not np.array_equal(dst, O.warp_perspective(img, M, (cw, ch))):
            fail("warp model", case=case, M=M.tolist(), src=[w, h], dst=[cw, ch])
        n["warp"] += 1
        n["warp_px"] += cw * ch
    print(json.dumps({"ok": True, "seconds": round(time.time() - t0, 1), "cases": a.cases, **n}))


if __name__ == "__main__":
    main()
