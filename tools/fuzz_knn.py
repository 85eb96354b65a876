"""CPU-only fuzz of the opt-in 2-NN / Lowe-ratio matcher (pano_match_knn):

    python tools/fuzz_knn.py --cases 300 --seed 1

Per case: a random image pair (smooth or quantised so that distance ties are common), random keypoints (some outside
the in-border region, some duplicated), a random ratio / patch size / descriptor.  Three implementations must agree:
  * the checker (oracle.match_knn) against real cv2.BFMatcher.knnMatch(k = 2) + Lowe's test: the same query set passes,
    with the same (nearest, runner-up) distances (cv2 leaves the order of ties unspecified, so indices are compared
    against the brute-force "earliest train keypoint" rule instead);
  * the engine's kernels (csrc/knn_kernels.cuh compiled unchanged on the CPU emulation of the CUDA execution model,
    tests/hostsim) against the checker, bit for bit, for a random split of the train range and block order;
  * the tensor-core epilogue's arithmetic (tests/hostsim emu_tc_top2) and the tensor-core kernel itself
    (match_tc_top2_kernel on the host model of mbarriers / TMA / tcgen05, tests/hostsim/tcgen05_emu.hpp) against the
    brute force on the same descriptors.
One JSON line; exit code 1 on the first difference."""
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
MATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("distance", "<f4")])


def p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def load_tc_emu():
    d = os.path.join(ROOT, "tests", "hostsim")
    so = os.path.join(d, "libmatch_tc_emu.so")
    srcs = [os.path.join(d, "match_tc_emu.cpp"), os.path.join(d, "cuda_emu.hpp"), os.path.join(d, "tcgen05_emu.hpp")]
    srcs += [os.path.join(ROOT, PKG, "csrc", f) for f in ("match_tc_kernels.cuh", "knn_core.cuh", "pano_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas",
                               "-Wno-unused-function", "-o", so, srcs[0]])
    lib = C.CDLL(so)
    lib.tcemu_last_error.restype = C.c_char_p
    return lib


def load_emu():
    d = os.path.join(ROOT, "tests", "hostsim")
    so = os.path.join(d, "libknn_emu.so")
    srcs = [os.path.join(d, "knn_emu.cpp"), os.path.join(d, "cuda_emu.hpp")]
    srcs += [os.path.join(ROOT, PKG, "csrc", f) for f in ("knn_kernels.cuh", "knn_core.cuh", "pano_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas",
                               "-o", so, srcs[0]])
    lib = C.CDLL(so)
    lib.emu_last_error.restype = C.c_char_p
    return lib


def random_image(rng, w, h):
    kind = rng.random()
    if kind < 0.4:
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    if kind < 0.7:     # few grey levels: many exact distance ties
        return (rng.integers(0, 4, (h, w, 3)) * 80).astype(np.uint8)
    base = rng.integers(0, 256, (h // 4 + 2, w // 4 + 2, 3)).astype(np.float64)
    up = np.kron(base, np.ones((4, 4, 1)))[:h, :w]
    return np.clip(up + rng.integers(-3, 4, (h, w, 3)), 0, 255).astype(np.uint8)


def random_keypoints(rng, w, h, n):
    k = np.stack([rng.integers(-1, w + 1, n), rng.integers(-1, h + 1, n)], 1).astype(np.int32)
    if n > 4 and rng.random() < 0.5:
        k[rng.integers(0, n, n // 5)] = k[rng.integers(0, n, n // 5)]      # duplicates
    return k


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=200)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    import cv2
    from oracle.oracle import Oracle
    O = Oracle()
    emu = load_emu()
    tc = load_tc_emu()
    rng = np.random.default_rng(a.seed)
    t0 = time.time()
    n = {"cases": 0, "queries": 0, "matches": 0, "pairs": 0, "tc_rows": 0}

    def fail(what, **kw):
        print(json.dumps({"ok": False, "difference": what, "fuzz_seed": a.seed, **kw}, default=str))
        sys.exit(1)
    for case in range(a.cases):
        w, h = int(rng.integers(8, 120)), int(rng.integers(8, 90))
        imq, imt = random_image(rng, w, h), random_image(rng, w, h)
        if rng.random() < 0.4:
            imt = imq.copy()
            imt[rng.integers(0, h):, :] = rng.integers(0, 256)
        kq = random_keypoints(rng, w, h, int(rng.integers(0, 300)))
        kt = random_keypoints(rng, w, h, int(rng.integers(0, 400)))
        descriptor = int(rng.integers(0, 2))
        patch = 5 if descriptor == 1 else int(rng.choice([1, 3, 5]))
        ratio = float(rng.choice([0.5, 0.6, 0.75, 0.8, 0.9, 0.99, 1.0]))
        b = patch // 2
        mo, so = O.match_knn(kq, kt, imq, imt, patch=patch, descriptor=descriptor, ratio=ratio)
        # ---- engine kernels on the CPU emulation --------------------------------------------------------------
        out = np.zeros(max(len(kq), 1), MATCH_DTYPE)
        sec = np.zeros(max(len(kq), 1), np.float32)
        kqc, ktc = np.ascontiguousarray(kq), np.ascontiguousarray(kt)
        m = emu.emu_match_knn(p(kqc, C.c_int32), len(kqc), p(ktc, C.c_int32), len(ktc), p(imq, C.c_uint8), w, h,
                              C.c_size_t(imq.strides[0]), p(imt, C.c_uint8), w, h, C.c_size_t(imt.strides[0]), patch, descriptor,
                              C.c_double(ratio), int(rng.integers(0, 9)), int(rng.integers(0, 3)), out.ctypes.data_as(C.c_void_p),
                              p(sec, C.c_float), len(out))
        if m != len(mo) or not np.array_equal(out[:m], mo) or not np.array_equal(sec[:m], so):
            fail("emulated kernels vs checker", case=case, n=(m, len(mo)), err=emu.emu_last_error())
        # ---- checker against cv2.BFMatcher -------------------------------------------------------------------
        inq = ~((kq[:, 0] < b) | (kq[:, 1] < b) | (kq[:, 0] + b >= w) | (kq[:, 1] + b >= h)) if len(kq) else np.zeros(0, bool)
        intr = ~((kt[:, 0] < b) | (kt[:, 1] < b) | (kt[:, 0] + b >= w) | (kt[:, 1] + b >= h)) if len(kt) else np.zeros(0, bool)
        qi, ti = np.flatnonzero(inq), np.flatnonzero(intr)
        n["cases"] += 1
        n["queries"] += len(qi)
        n["matches"] += len(mo)
        n["pairs"] += len(qi) * len(ti)
        if len(qi) == 0 or len(ti) < 2:
            if len(mo):
                fail("matches without two candidates", case=case)
            continue
        if descriptor == 0:
            dq = np.stack([imq[y - b:y + b + 1, x - b:x + b + 1].reshape(-1) for x, y in kq[qi]]).astype(np.float32)
            dt = np.stack([imt[y - b:y + b + 1, x - b:x + b + 1].reshape(-1) for x, y in kt[ti]]).astype(np.float32)
            bf, factor = cv2.BFMatcher(cv2.NORM_L2SQR), ratio * ratio
            D = ((dq[:, None, :] - dt[None, :, :]) ** 2).sum(-1).astype(np.int64)
        else:
            dq = np.stack([O.knn_binary_descriptor(imq, x, y) for x, y in kq[qi]]).view(np.uint8)
            dt = np.stack([O.knn_binary_descriptor(imt, x, y) for x, y in kt[ti]]).view(np.uint8)
            bf, factor = cv2.BFMatcher(cv2.NORM_HAMMING), ratio
            D = np.unpackbits(dq[:, None, :] ^ dt[None, :, :], axis=-1).sum(-1).astype(np.int64)
        expect = {}
        for r, (m1, m2) in enumerate(bf.knnMatch(dq, dt, k=2)):
            d1, d2 = int(m1.distance), int(m2.distance)
            if float(d1) < factor * float(d2):
                expect[int(qi[r])] = (d1, d2)
        if sorted(expect) != [int(q) for q in mo["queryIdx"]]:
            fail("checker vs cv2: passing query set", case=case)
        row_of = {int(q): r for r, q in enumerate(qi)}
        for rec, s in zip(mo, so):
            q = int(rec["queryIdx"])
            if (int(rec["distance"]), int(s)) != expect[q]:
                fail("checker vs cv2: distances", case=case, q=q)
            row = D[row_of[q]]
            if int(rec["trainIdx"]) != int(ti[np.flatnonzero(row == expect[q][0])[0]]):
                fail("checker: not the earliest train keypoint at the nearest distance", case=case, q=q)
        # ---- tensor-core epilogue arithmetic on the patch descriptors ---------------------------------------------
        if descriptor == 0 and len(qi) and len(ti):
            qd = np.zeros((len(qi), 128), np.uint8)
            td = np.zeros((len(ti), 128), np.uint8)
            qd[:, :dq.shape[1]] = dq.astype(np.uint8)
            td[:, :dt.shape[1]] = dt.astype(np.uint8)
            b1 = np.zeros(len(qi), np.uint64)
            b2 = np.zeros(len(qi), np.uint64)
            emu.emu_tc_top2(p(qd, C.c_uint8), len(qi), p(td, C.c_uint8), len(ti), int(rng.integers(1, 5)), int(rng.integers(0, 3)),
                            p(b1, C.c_uint64), p(b2, C.c_uint64))
            order = np.lexsort((np.broadcast_to(np.arange(len(ti)), D.shape), D), axis=1)
            for r in range(len(qi)):
                j1, j2 = int(order[r, 0]), int(order[r, 1])
                if int(b1[r]) != (int(D[r, j1]) << 32 | j1) or int(b2[r]) != (int(D[r, j2]) << 32 | j2):
                    fail("tensor-core epilogue arithmetic", case=case, row=r)
            n["tc_rows"] += len(qi)
            # ... and the tensor-core kernel itself (match_tc_top2_kernel on the host model of tcgen05 / TMA / mbarriers)
            c1 = np.zeros(len(qi), np.uint64)
            c2 = np.zeros(len(qi), np.uint64)
            st = tc.tcemu_match(p(qd, C.c_uint8), len(qi), p(td, C.c_uint8), len(ti), int(rng.choice([1, 2, 5, 148])), 1,
                                int(rng.integers(0, 3)), p(c1, C.c_uint64), p(c2, C.c_uint64))
            if st != 0 or not np.array_equal(c1, b1) or not np.array_equal(c2, b2):
                fail("tensor-core top-2 kernel on the host model", case=case, st=st, err=tc.tcemu_last_error())
    print(json.dumps({"ok": True, "fuzz_seed": a.seed, **n, "seconds": round(time.time() - t0, 1), "cv2": cv2.__version__}))


if __name__ == "__main__":
    main()
