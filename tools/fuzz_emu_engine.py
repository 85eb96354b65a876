"""CPU-only fuzz of the WHOLE engine against the reference semantics:

    python tools/fuzz_emu_engine.py --cases 200 --seed 1

tools/fuzz_emu_kernels.py drives the kernels one by one; this tool drives the LIBRARY - the C ABI, the host orchestration
(stage sequencing, scratch buffers, read-backs, replay planning / pre-launch / re-runs, the async worker, fold and batch
loops) and every kernel - as tests/hostsim/build_emu_lib.py builds it from the engine's own sources on the CPU emulation
of the CUDA execution model, through the package's Python binding, exactly as a caller on a B200 would.  Per case: a
random scene (the kernel fuzz's generator: textured pairs, noise, quantised images with exact SSD ties, flat images,
sparse blobs, odd and tiny sizes), random detector / matcher / RANSAC options, random seed, random engine settings
(tensor-core or SIMT matcher, chunked or resident replay, the reference's matcher or the opt-in ratio-test matcher,
blocking or asynchronous call) and one of: the stage entry points one by one, fused pair, homography only, three-image fold (the reference's or the opt-in
incremental one is left to its own test), batch of pairs.  Everything the call returns - status, counts, best inlier
count, H bits, canvas bytes - must equal the composition of the oracle's stage functions (detect, match / match_knn,
ransac, compose; ref: src/serial/main.cpp:311-414).  One JSON line; exit code 1 on the first difference.
TEST INFRASTRUCTURE: the emulated library is never shipped."""
import argparse
import ctypes as C
import importlib
import importlib.util
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "ucb-cs267-parallel-panoramic-image-stitching_b200"
MAX_CANVAS_PX = 400_000        # larger canvases: the homography-only call (the emulation renders ~1 MP/s)


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def expected(oracle, left, right, ho, ro, seed, knn):
    """the reference's stitchTwoImages as a composition of the checker's stage functions -> (status name, fields)"""
    kl = oracle.detect(left, k=ho.k_, thresh=ho.nmsThresh_, nbhd=ho.nmsNeighborhood_)
    kr = oracle.detect(right, k=ho.k_, thresh=ho.nmsThresh_, nbhd=ho.nmsNeighborhood_)
    if knn is None:
        m = oracle.match(kr, kl, right, left, patch=ho.patchSize_, max_ssd=ho.maxSSDThresh_)
    else:
        m, _ = oracle.match_knn(kr, kl, right, left, patch=ho.patchSize_, descriptor=knn[1], ratio=knn[0])
    e = dict(kl=len(kl), kr=len(kr), m=len(m), H=None, best=None, canvas=None, geom=None)
    if len(m) == 0:
        return "NO_MATCHES", e
    if len(m) < 4:
        return "TOO_FEW_MATCHES", e
    o = oracle.ransac(kr, kl, m, iters=ro.numIterations_, thr=ro.distanceThreshold_, seed=seed)
    if not o["ok"]:
        return "NO_HOMOGRAPHY", e
    e["H"], e["best"] = o["H"], o["best_count"]
    ok, geom, _ = oracle.canvas_geometry(left.shape[1], left.shape[0], right.shape[1], right.shape[0], o["H"])
    e["geom"] = geom if ok else None
    if not ok:
        return "ROI", e
    return "OK", e


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=100)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    os.environ["PANO_BATCH_LANES"] = "1"          # the emulation runs one launch at a time
    pkg = importlib.import_module(PKG)
    synth = importlib.import_module(PKG + ".synth")
    from oracle.oracle import Oracle
    oracle = Oracle()
    kfuzz = _load(os.path.join(ROOT, "tools", "fuzz_emu_kernels.py"), "fuzz_emu_kernels")
    emu = _load(os.path.join(ROOT, "tests", "hostsim", "build_emu_lib.py"), "build_emu_lib")
    lib = C.CDLL(emu.build())
    lib.pano_last_error.restype = C.c_char_p
    lib.pano_version.restype = C.c_char_p
    lib.pano_kernel_launches.restype = C.c_uint64
    eng = pkg.Engine.__new__(pkg.Engine)
    eng.lib, eng.ctx, eng.device = lib, C.c_void_p(), 0
    assert lib.pano_create(0, C.c_uint32(12345), C.byref(eng.ctx)) == 0
    names = {v: k for k, v in vars(pkg).items() if k.startswith("PANO_ERR_") or k == "PANO_OK"}

    def status_name(code):
        return names.get(code, str(code)).replace("PANO_ERR_", "").replace("PANO_OK", "OK")

    rng = np.random.default_rng(a.seed)
    tot = dict(cases=0, stage_calls=0, pair=0, homography_only=0, fold=0, batch=0, asynchronous=0, ratio_test_matcher=0, resident_replay=0,
               simt_matcher=0, status_ok=0, no_matches=0, too_few=0, no_homography=0, roi=0, canvas_px=0, ransac_iterations=0)
    t0 = time.time()

    def fail(case, what, **kw):
        print(json.dumps(dict(ok=False, fuzz_seed=a.seed, case=case, what=what, **{k: str(v) for k, v in kw.items()})))
        sys.exit(1)

    def check_record(case, tag, r, st, e):
        if status_name(r["status"]) != st:
            fail(case, tag + ": status", got=status_name(r["status"]), want=st, err=lib.pano_last_error(eng.ctx))
        if (r["kl"], r["kr"], r["m"]) != (e["kl"], e["kr"], e["m"]):
            fail(case, tag + ": counts", got=(r["kl"], r["kr"], r["m"]), want=(e["kl"], e["kr"], e["m"]))
        if e["H"] is not None and (r["best"] != e["best"] or not np.array_equal(bits(r["H"]), bits(e["H"]))):
            fail(case, tag + ": homography", got=(r["best"], r["H"].tolist()), want=(e["best"], e["H"].tolist()))
        if st == "OK" and tuple(r["canvas"]) != tuple(e["geom"]):
            fail(case, tag + ": canvas geometry", got=r["canvas"], want=e["geom"])
        key = {"OK": "status_ok", "NO_MATCHES": "no_matches", "TOO_FEW_MATCHES": "too_few", "NO_HOMOGRAPHY": "no_homography", "ROI": "roi"}[st]
        tot[key] += 1

    for case in range(a.cases):
        left, right = kfuzz.scene(rng, synth)
        ho = pkg.HarrisCornerOptions(k_=0.04, nmsThresh_=float(rng.choice([1e6, 1e6, 1e5, 3e6, 1e4])),
                                     nmsNeighborhood_=int(rng.choice([3, 3, 5, 7])), patchSize_=int(rng.choice([5, 5, 5, 3, 1])),
                                     maxSSDThresh_=float(rng.choice([1e8, 1e8, 1e8, 20000.0, 3000.0])))
        ro = pkg.RansacOptions(numIterations_=int(rng.integers(3, 48)), distanceThreshold_=float(rng.choice([3.0, 3.0, 1.0, 6.0])))
        seed = int(rng.integers(0, 1 << 32))
        knn = None
        if rng.random() < 0.25:
            knn = (float(rng.choice([0.6, 0.75, 0.9, 1.0])), int(rng.random() < 0.3 and ho.patchSize_ == 5))
            tot["ratio_test_matcher"] += 1
        simt, resident = bool(rng.random() < 0.35), bool(rng.random() < 0.3)
        tot["simt_matcher"] += simt
        tot["resident_replay"] += resident
        eng.set_seed(seed)
        eng.set_matcher(1 if simt else 0)
        eng.set_replay_mode(1 if resident else 0)
        eng.set_match_mode(1 if knn else 0, *(knn or (0.75, 0)))
        tot["cases"] += 1
        tot["ransac_iterations"] += ro.numIterations_
        op = rng.random()
        try:
            if op < 0.15:
                # the stage entry points (ref: src/gpu/*.cuh), blocking or asynchronous form, one by one
                eng.set_match_mode(0)
                use_async = bool(rng.random() < 0.5)
                tot["stage_calls"] += 1
                tot["asynchronous"] += use_async
                hk = dict(k=ho.k_, nmsThresh=ho.nmsThresh_, nmsNeighborhood=ho.nmsNeighborhood_)
                if use_async:
                    kl = eng.gpuHarrisCornerDetectorDetectAsync(left, **hk).result()
                    kr = eng.gpuHarrisCornerDetectorDetectAsync(right, **hk).result()
                else:
                    kl, kr = eng.gpuHarrisCornerDetectorDetect(left, **hk), eng.gpuHarrisCornerDetectorDetect(right, **hk)
                okl = oracle.detect(left, k=ho.k_, thresh=ho.nmsThresh_, nbhd=ho.nmsNeighborhood_)
                okr = oracle.detect(right, k=ho.k_, thresh=ho.nmsThresh_, nbhd=ho.nmsNeighborhood_)
                if not (np.array_equal(kl, okl) and np.array_equal(kr, okr)):
                    fail(case, "detect", got=(len(kl), len(kr)), want=(len(okl), len(okr)))
                off = int(rng.choice([0, 0, 3]))
                mk = dict(patchSize=ho.patchSize_, maxSSDThresh=ho.maxSSDThresh_, offset=off)
                m = (eng.gpuHarrisMatchKeyPointsAsync(kr, kl, right, left, **mk).result() if use_async
                     else eng.gpuHarrisMatchKeyPoints(kr, kl, right, left, **mk))
                om = np.ascontiguousarray(oracle.match(okr, okl, right, left, patch=ho.patchSize_, max_ssd=ho.maxSSDThresh_, offset=off))
                if np.ascontiguousarray(m).tobytes() != om.tobytes():
                    fail(case, "match", got=len(m), want=len(om))
                if off == 0 and len(om) > 0:
                    o = oracle.ransac(okr, okl, om, iters=ro.numIterations_, thr=ro.distanceThreshold_, seed=seed)
                    if use_async:
                        H, best, _ = eng.computeHomographyAsync(kr, kl, m, options=ro).result()
                        same = (H is None) == (not o["ok"]) and (H is None or (best == o["best_count"] and np.array_equal(bits(H), bits(o["H"]))))
                    else:
                        d = eng.computeHomography(kr, kl, m, options=ro, details=True)
                        same = d["ok"] == o["ok"] and (len(om) < 4 or (np.array_equal(d["samples"], o["samples"]) and np.array_equal(d["counts"], o["counts"])))
                        same = same and (not o["ok"] or (np.array_equal(bits(d["H"]), bits(o["H"])) and d["best_count"] == o["best_count"]
                                                          and np.array_equal(d["inlier_mask"], o["inlier_mask"])))
                    if not same:
                        fail(case, "ransac", matches=len(om), oracle_ok=o["ok"])
                    if o["ok"]:
                        ok, geom, _ = oracle.canvas_geometry(left.shape[1], left.shape[0], right.shape[1], right.shape[0], o["H"])
                        if ok and geom[0] * geom[1] <= MAX_CANVAS_PX:
                            if not np.array_equal(eng.warpOverlay(left, right, o["H"]), oracle.compose(left, right, o["H"])):
                                fail(case, "warp_overlay")
                            tot["canvas_px"] += geom[0] * geom[1]
            elif op < 0.7:
                st, e = expected(oracle, left, right, ho, ro, seed, knn)
                big = st == "OK" and e["geom"][0] * e["geom"][1] > MAX_CANVAS_PX
                if big or rng.random() < 0.15:
                    r = eng.pairHomography(left, right, harrisOpts=ho, ransacOpts=ro)
                    tot["homography_only"] += 1
                    if st == "ROI":                  # (no canvas is composed by this call: the geometry is not judged)
                        st = "OK"
                        e["geom"] = tuple(r["canvas"])
                    check_record(case, "pair_homography", r, st, e)
                else:
                    if rng.random() < 0.3:
                        tot["asynchronous"] += 1
                        h = eng.stitchTwoImagesAsync(left, right, harrisOpts=ho, ransacOpts=ro)
                        r = h.result()
                        canvas = eng.getCanvas() if r["status"] == 0 else None
                    else:
                        canvas, r = eng.stitchTwoImages(left, right, harrisOpts=ho, ransacOpts=ro)
                    tot["pair"] += 1
                    check_record(case, "stitch_pair", r, st, e)
                    if st == "OK":
                        want = oracle.compose(left, right, e["H"])
                        if not np.array_equal(canvas, want):
                            fail(case, "stitch_pair: canvas bytes", shape=(None if canvas is None else canvas.shape), want=want.shape)
                        tot["canvas_px"] += want.shape[0] * want.shape[1]
            elif op < 0.85:
                # three-image fold (ref: stitchAllImages): step 2 detects on the panorama of step 1
                third = np.ascontiguousarray(np.roll(right, int(rng.integers(2, 12)), axis=1))
                views = [left, right, third]
                want, steps, stop = np.ascontiguousarray(left), [], None
                for im in views[1:]:
                    st, e = expected(oracle, want, im, ho, ro, seed, knn)
                    steps.append((st, e))
                    if st != "OK":
                        stop = st
                        break
                    nxt = oracle.compose(want, im, e["H"])
                    if nxt.shape[0] * nxt.shape[1] > MAX_CANVAS_PX:
                        stop = "too large for the emulation"
                        break
                    want = nxt
                if stop is None:
                    pano, log = eng.stitchAllImages(views, harrisOpts=ho, ransacOpts=ro)
                    for i, (st, e) in enumerate(steps):
                        check_record(case, "fold step %d" % i, log[i], st, e)
                    if not np.array_equal(np.asarray(pano), want):
                        fail(case, "fold: panorama bytes")
                    tot["fold"] += 1
                    tot["canvas_px"] += want.shape[0] * want.shape[1]
            else:
                # batch of pairs of one geometry (throughput mode's host loop: stage A / stage B slots)
                n = int(rng.integers(2, 4))
                lefts = [left] + [np.ascontiguousarray(np.roll(left, int(rng.integers(1, 9)), axis=0)) for _ in range(n - 1)]
                rights = [right] + [np.ascontiguousarray(np.roll(right, int(rng.integers(1, 9)), axis=0)) for _ in range(n - 1)]
                exp = [expected(oracle, l, r, ho, ro, seed, knn) for l, r in zip(lefts, rights)]
                if all(not (st == "OK" and e["geom"][0] * e["geom"][1] > MAX_CANVAS_PX) for st, e in exp):
                    cap = max([3 * e["geom"][0] * e["geom"][1] for st, e in exp if st == "OK"] + [64]) + 256
                    outs = [np.zeros(cap, np.uint8) for _ in range(n)]
                    res, _ = eng.stitchBatch(lefts, rights, harrisOpts=ho, ransacOpts=ro, canvases_out=outs)
                    for i, (st, e) in enumerate(exp):
                        check_record(case, "batch pair %d" % i, res[i], st, e)
                        if st == "OK":
                            cw, ch = e["geom"][0], e["geom"][1]
                            want = oracle.compose(lefts[i], rights[i], e["H"])
                            if not np.array_equal(outs[i][:3 * cw * ch].reshape(ch, cw, 3), want):
                                fail(case, "batch pair %d: canvas bytes" % i)
                            tot["canvas_px"] += cw * ch
                    tot["batch"] += 1
        except pkg.PanoError as ex:
            fail(case, "engine error", err=ex)
    eng.set_match_mode(0)
    eng.close()
    print(json.dumps(dict(ok=True, fuzz_seed=a.seed, **{k: int(v) for k, v in tot.items()}, seconds=round(time.time() - t0, 1))))


if __name__ == "__main__":
    main()
