"""CPU-only fuzz of the engine's own kernel source against the reference semantics:

    python tools/fuzz_emu_kernels.py --cases 300 --seed 1

The device code of the four stages (csrc/harris_kernels.cuh, match_kernels.cuh, ransac_kernels.cuh, warp_kernels.cuh) is
compiled unchanged by g++ on the CPU emulation of the CUDA execution model (tests/hostsim/cuda_emu.hpp) and driven
through a whole pair per case - detect both images (a random one of: fused kernel with plain loads, fused kernel with
the modelled TMA load, two-kernel path), match (random split of the train range), RANSAC (chunked or resident replay,
random chunk target, few iterations), warp + overlay (quad / round-1 fast / general kernel) - on random scenes (textured
pairs, noise, flat, quantised images with exact SSD ties, odd and tiny sizes) and compared stage by stage with the
oracle (keypoints, match records, samples, inlier counts, H bits, canvas bytes).  One JSON line; exit code 1 on the
first difference."""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "ucb-cs267-parallel-panoramic-image-stitching_b200"
MATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("distance", "<f4")])


def p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def load(name, headers):
    d = os.path.join(ROOT, "tests", "hostsim")
    so = os.path.join(d, "lib%s.so" % name)
    srcs = [os.path.join(d, name + ".cpp"), os.path.join(d, "cuda_emu.hpp")]
    srcs += [os.path.join(ROOT, PKG, "csrc", f) for f in headers]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas",
                               "-Wno-unused-function", "-o", so, srcs[0]])
    return C.CDLL(so)


def scene(rng, synth):
    kind = rng.random()
    w, h = int(rng.integers(24, 200)), int(rng.integers(24, 150))
    if kind < 0.45:
        left, right, _ = synth.make_pair(w + 40, h + 30, seed=int(rng.integers(0, 1 << 30)), rot_deg=float(rng.uniform(0, 0.5)))
        return np.ascontiguousarray(left[:h, :w]), np.ascontiguousarray(right[:h + int(rng.integers(0, 20)), :w + int(rng.integers(0, 30))])
    if kind < 0.65:
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        b = np.roll(a, int(rng.integers(1, w // 2)), axis=1)           # shifted copy: real correspondences
        return a, np.ascontiguousarray(b)
    if kind < 0.8:
        a = (rng.integers(0, 3, (h, w, 3)) * 120).astype(np.uint8)      # quantised: exact SSD ties
        return a, np.ascontiguousarray(a[:, ::-1])
    if kind < 0.9:
        return np.full((h, w, 3), int(rng.integers(0, 256)), np.uint8), rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    a = np.zeros((h, w, 3), np.uint8)
    for _ in range(int(rng.integers(3, 30))):                           # sparse blobs on black
        x, y = int(rng.integers(0, w - 4)), int(rng.integers(0, h - 4))
        a[y:y + int(rng.integers(2, 9)), x:x + int(rng.integers(2, 9))] = rng.integers(60, 256, 3)
    return a, np.ascontiguousarray(np.roll(a, (int(rng.integers(0, 6)), int(rng.integers(2, 20))), axis=(0, 1)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=100)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    from oracle.oracle import Oracle
    O = Oracle()
    synth = importlib.import_module(PKG + ".synth")
    H_ = load("harris_emu", ("harris_kernels.cuh", "pano_core.cuh"))
    M_ = load("match_emu", ("match_kernels.cuh", "harris_kernels.cuh", "pano_core.cuh"))
    R_ = load("ransac_emu", ("ransac_kernels.cuh", "replay_plan.hpp", "pano_core.cuh"))
    W_ = load("warp_emu", ("warp_kernels.cuh", "pano_core.cuh"))
    rng = np.random.default_rng(a.seed)
    t0 = time.time()
    n = {"cases": 0, "keypoints": 0, "matches": 0, "ransac_iterations": 0, "homographies": 0, "canvas_px": 0}

    def fail(what, **kw):
        print(json.dumps({"ok": False, "difference": what, "fuzz_seed": a.seed, **kw}, default=str))
        sys.exit(1)

    def detect(img, thresh, nbhd, path):
        h, w = img.shape[:2]
        xy = np.zeros((w * h // 2 + 16, 2), np.int32)
        k = H_.hemu_detect(p(img, C.c_uint8), w, h, C.c_size_t(img.strides[0]), C.c_double(0.04), C.c_double(thresh), nbhd, path,
                           int(rng.integers(0, 3)), p(xy, C.c_int32), len(xy))
        if k < 0:
            fail("detect: emulation error", k=k)
        return xy[:k]

    for case in range(a.cases):
        left, right = scene(rng, synth)
        thresh = float(rng.choice([1e6, 1e5, 1e4]))
        nbhd = 3 if rng.random() < 0.8 else int(rng.choice([5, 7]))
        path = int(rng.integers(0, 2)) if nbhd == 3 and rng.random() < 0.8 else 2
        kl, kr = detect(left, thresh, nbhd, path), detect(right, thresh, nbhd, path)
        kol, kor = O.detect(left, thresh=thresh, nbhd=nbhd), O.detect(right, thresh=thresh, nbhd=nbhd)
        if not (np.array_equal(kl, kol) and np.array_equal(kr, kor)):
            fail("keypoints", case=case, path=path, nbhd=nbhd)
        n["cases"] += 1
        n["keypoints"] += len(kl) + len(kr)
        # ---- match (query = right, train = left, as the reference calls it) ---------------------------------------
        patch = int(rng.choice([5, 5, 5, 3, 1]))
        max_ssd = float(rng.choice([1e8, 1e8, 5000.0, 800.0]))
        out = np.zeros(max(len(kr), 1), MATCH_DTYPE)
        m = M_.memu_match(p(np.ascontiguousarray(kr), C.c_int32), len(kr), p(np.ascontiguousarray(kl), C.c_int32), len(kl),
                          p(right, C.c_uint8), right.shape[1], right.shape[0], C.c_size_t(right.strides[0]),
                          p(left, C.c_uint8), left.shape[1], left.shape[0], C.c_size_t(left.strides[0]), patch, C.c_double(max_ssd),
                          0, int(rng.integers(0, 6)), out.ctypes.data_as(C.c_void_p), len(out), None)
        mo = np.ascontiguousarray(O.match(kr, kl, right, left, patch=patch, max_ssd=max_ssd))
        if m != len(mo) or out[:m].tobytes() != mo.tobytes():
            fail("matches", case=case, n=(m, len(mo)))
        n["matches"] += m
        if m < 4:
            continue
        # ---- RANSAC ---------------------------------------------------------------------------------------------------
        iters = int(rng.integers(5, 40))
        seed = int(rng.integers(0, 1 << 31))
        thr = float(rng.choice([3.0, 3.0, 1.0, 6.0]))
        o = O.ransac(kr, kl, mo, iters=iters, thr=thr, seed=seed)
        Hm = np.zeros((3, 3))
        best, best_iter, errw, chunks = C.c_int(0), C.c_int(-1), C.c_int(0), C.c_int(0)
        samples = np.full((iters, 4), -1, np.int32)
        counts = np.full(iters, -2, np.int32)
        mask = np.zeros(m, np.uint8)
        st = R_.remu_ransac(p(np.ascontiguousarray(kr), C.c_int32), len(kr), p(np.ascontiguousarray(kl), C.c_int32), len(kl),
                            mo.ctypes.data_as(C.c_void_p), m, iters, C.c_double(thr), C.c_uint32(seed), int(rng.integers(0, 2)),
                            C.c_double(float(rng.choice([50000.0, 4000.0, 300.0]))), C.c_double(4.2), 1, p(Hm, C.c_double),
                            C.byref(best), C.byref(best_iter), p(samples, C.c_int32), p(counts, C.c_int32), p(mask, C.c_uint8),
                            C.byref(errw), C.byref(chunks))
        if st != (0 if o["ok"] else 5):
            fail("ransac status", case=case, st=st, ok=o["ok"])
        if not (np.array_equal(samples, o["samples"][:iters]) and np.array_equal(counts, o["counts"][:iters])):
            fail("ransac samples / counts", case=case)
        if (best.value, best_iter.value) != (o["best_count"], o["best_iter"]):
            fail("ransac best", case=case)
        n["ransac_iterations"] += iters
        if not o["ok"]:
            continue
        if not (np.array_equal(Hm.view(np.uint64), np.ascontiguousarray(o["H"]).view(np.uint64))
                and np.array_equal(mask.astype(bool), o["inlier_mask"])):
            fail("homography / inlier mask", case=case)
        n["homographies"] += 1
        # ---- warp + overlay -------------------------------------------------------------------------------------------
        cap = 48 << 20
        okg, (gw, gh, _, _), _ = O.canvas_geometry(left.shape[1], left.shape[0], right.shape[1], right.shape[0], o["H"])
        if okg and gw * gh * 3 > cap:      # a wild homography of a degenerate scene: canvas too large to compare
            continue
        want = O.compose(left, right, o["H"])
        canvas = np.zeros(cap if want is not None else 16, np.uint8)
        geom = (C.c_int * 5)()
        st = W_.wemu_overlay(p(left, C.c_uint8), left.shape[1], left.shape[0], C.c_size_t(left.strides[0]),
                             p(right, C.c_uint8), right.shape[1], right.shape[0], C.c_size_t(right.strides[0]),
                             p(np.ascontiguousarray(o["H"]), C.c_double), int(rng.integers(0, 3)), int(rng.choice([256, 256, 4, 1])),
                             p(canvas, C.c_uint8), C.c_size_t(len(canvas)), geom)
        if want is None:
            if st not in (0, -2):
                fail("canvas geometry", case=case, st=st)
            continue
        if st != 1 or not np.array_equal(canvas[:want.size].reshape(want.shape), want):
            fail("canvas", case=case, st=st)
        n["canvas_px"] += want.shape[0] * want.shape[1]
    print(json.dumps({"ok": True, "fuzz_seed": a.seed, **n, "seconds": round(time.time() - t0, 1)}))


if __name__ == "__main__":
    main()
