"""CPU tier: what PINS the oracle.

(1) oracle/_ref is the reference ITSELF — /root/reference/src/serial/main.cpp compiled unmodified
    against oracle/cvshim (cv2-pinned OpenCV arithmetic) with std::random_device replaced by a
    fixed seed.  Where it is available (built here, prebuilt on the GPU box) the oracle
    restatement (oracle/pano_oracle.cpp) is held to it stage by stage, live.
(2) tests/golden/ref_small.npz and ref_runs.json are outputs of that reference run in the build
    container (oracle/gen_ref_golden.py); the oracle is held to them everywhere.
(3) SURVEY §8 c2-(ii): the oracle's restated OpenCV pieces against the REAL cv2 routines at full
    size on the benchmark pair (warpPerspective + overlay on the 4K canvas, all 1000
    findHomography calls of a RANSAC run, gray, gemm 3x1).
"""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_synth


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def hexbits(H):
    return [format(int(v), "016x") for v in bits(H).ravel()]


@pytest.fixture(scope="module")
def ref():
    from oracle import ref as refmod
    if not refmod.available() and not os.path.exists(os.path.join(refmod.REFERENCE_ROOT, "src", "serial", "main.cpp")):
        pytest.skip("oracle/_ref not built and the reference sources are not present")
    return refmod.Reference()


@pytest.fixture(scope="module")
def small():
    return np.load(os.path.join(GOLDEN, "ref_small.npz"))


@pytest.fixture(scope="module")
def runs():
    f = os.path.join(GOLDEN, "ref_runs.json")
    if not os.path.exists(f):
        pytest.skip("tests/golden/ref_runs.json not generated (oracle/gen_ref_golden.py)")
    return json.load(open(f))


# ---------------- (1) the recipe ------------------------------------------------------------
def test_ref_recipe_compiles_the_reference_sources_in_place():
    """oracle/Makefile names the reference's files where they lie; nothing of them is in the repo"""
    mk = open(os.path.join(ROOT, "oracle", "Makefile")).read()
    assert "$(REF)/src/reader/reader.cpp" in mk and "$(REF)/src/serial/main.cpp" in mk
    bridge = open(os.path.join(ROOT, "oracle", "ref_bridge.cpp")).read()
    assert '#include "serial/main.cpp"' in bridge and '#include "openmp/main.cpp"' in bridge
    for dirpath, _, files in os.walk(ROOT):
        if "/.git" in dirpath or "/gpurun_out" in dirpath:
            continue
        for f in files:
            if f.endswith((".cpp", ".cu", ".hpp", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                # a line only the reference has (src/serial/main.cpp:380): would betray a copied source
                assert "Simple overlay fusion: for non-black pixels" not in txt, f


def test_ref_golden_files_are_current(ref, small):
    """the committed vectors are what the reference produces now (guards stale fixtures)"""
    synth = load_synth()
    for tag in ("a", "b"):
        w, h, s = (int(v) for v in small["%s_size" % tag])
        left, right, _ = synth.make_pair(w, h, seed=s)
        assert np.array_equal(ref.detect(left), small["%s_kl" % tag])
        r = ref.stitch_pair(left, right, seed=12345)
        assert r["status"] == 1 and np.array_equal(r["canvas"], small["%s_canvas" % tag])
        assert set(r["times_ms"]) >= {"Harris Corner Detection", "Harris Corner Matching",
                                      "RANSAC Homography Estimation", "Image Stitching"}


# ---------------- (1) oracle == reference, live, stage by stage --------------------------------
@pytest.mark.parametrize("w,h,seed", [(480, 270, 11), (333, 201, 5), (640, 360, 3), (257, 403, 8)])
def test_oracle_equals_reference_stage_by_stage(ref, oracle, w, h, seed):
    left, right, _ = load_synth().make_pair(w, h, seed=seed)
    kl, kr = oracle.detect(left), oracle.detect(right)
    assert np.array_equal(kl, ref.detect(left)) and np.array_equal(kr, ref.detect(right))
    m = oracle.match(kr, kl, right, left)
    mr = ref.match(kr, kl, right, left)
    assert len(m) > 20 and np.array_equal(m, mr)
    for rs in (12345, 1, 267):
        ro = oracle.ransac(kr, kl, m, seed=rs)
        Hr = ref.ransac(kr, kl, m, seed=rs)
        assert ro["ok"] == (Hr is not None)
        if ro["ok"]:
            assert np.array_equal(bits(ro["H"]), bits(Hr))
    so, sr = oracle.stitch_pair(left, right, seed=12345), ref.stitch_pair(left, right, seed=12345)
    assert (so["status"] == 1) == (sr["status"] == 1)
    if so["status"] == 1:
        assert np.array_equal(so["canvas"], sr["canvas"])


def test_oracle_equals_reference_convolution_and_taps(ref, oracle):
    rng = np.random.default_rng(4)
    plane = rng.integers(-1000, 1000, (37, 53)).astype(np.float64)
    assert np.array_equal(bits(oracle.gaussian_kernel(5, 1.0)), bits(ref.gaussian_kernel(5, 1.0)))
    for ks in (3, 5, 7):
        kern = rng.standard_normal((ks, ks))
        assert np.array_equal(bits(oracle.convolve(plane, kern)), bits(ref.convolve(plane, kern)))


def test_oracle_equals_reference_other_options(ref, oracle):
    left, right, _ = load_synth().make_pair(400, 300, seed=21)
    for k, thr, nb in ((0.04, 1e6, 5), (0.06, 1e5, 3), (0.04, 1e7, 7), (0.04, 0.0, 1)):
        assert np.array_equal(oracle.detect(left, k, thr, nb), ref.detect(left, k, thr, nb))
    kl, kr = oracle.detect(left), oracle.detect(right)
    for patch, mx, off in ((3, 1e8, 0), (5, 2000.0, 0), (5, 1e8, 7), (1, 1e8, 0)):
        assert np.array_equal(oracle.match(kr, kl, right, left, patch, mx, off),
                              ref.match(kr, kl, right, left, patch, mx, off))
    m = oracle.match(kr, kl, right, left)
    for iters, thr in ((50, 3.0), (200, 1.0), (1000, 10.0)):
        ro = oracle.ransac(kr, kl, m, iters=iters, thr=thr, seed=5)
        Hr = ref.ransac(kr, kl, m, iters=iters, thr=thr, seed=5)
        assert ro["ok"] == (Hr is not None) and (not ro["ok"] or np.array_equal(bits(ro["H"]), bits(Hr)))


def test_oracle_equals_reference_edge_cases(ref, oracle):
    flat = np.full((64, 96, 3), 77, np.uint8)
    assert len(ref.detect(flat)) == 0 == len(oracle.detect(flat))
    assert ref.stitch_pair(flat, flat)["status"] == 0 and oracle.stitch_pair(flat, flat)["status"] == 0
    # periodic texture: exact response ties must be rejected by both (strict NMS, ref :164-176)
    img = np.zeros((96, 96, 3), np.uint8)
    img[::8, :, :] = 255
    img[:, ::8, :] = 255
    assert np.array_equal(ref.detect(img), oracle.detect(img))
    # fewer matches than samples: the reference breaks out of the loop and returns an empty Mat (:268-269)
    kp = np.array([[10, 10], [20, 12], [30, 40]], np.int32)
    m = np.zeros(3, dtype=[("queryIdx", "<i4"), ("trainIdx", "<i4"), ("distance", "<f4")])
    m["queryIdx"] = m["trainIdx"] = np.arange(3)
    assert ref.ransac(kp, kp, m) is None and not oracle.ransac(kp, kp, m)["ok"]
    # degenerate samples (all x equal): findHomography returns an empty Mat every iteration
    kq = np.stack([np.full(8, 5), np.arange(8) * 7], 1).astype(np.int32)
    m8 = np.zeros(8, dtype=m.dtype)
    m8["queryIdx"] = m8["trainIdx"] = np.arange(8)
    assert ref.ransac(kq, kq, m8) is None and not oracle.ransac(kq, kq, m8)["ok"]
    # keypoints on the image border are skipped by the matcher on both sides (:204-207, :214-217)
    rng = np.random.default_rng(2)
    a, b = (rng.integers(0, 256, (40, 50, 3), dtype=np.uint8) for _ in range(2))
    k1 = np.array([[0, 0], [1, 5], [2, 2], [47, 37], [48, 20], [25, 38], [25, 20]], np.int32)
    assert np.array_equal(ref.match(k1, k1[::-1].copy(), a, b), oracle.match(k1, k1[::-1].copy(), a, b))


@pytest.mark.parametrize("m_count,iters", [(4, 30), (5, 30), (4097, 20), (65535, 6), (65536, 6), (70000, 6)])
def test_oracle_equals_reference_ransac_shuffle_branches(ref, oracle, m_count, iters):
    """both libstdc++ std::shuffle branches (paired draws up to 65 535 elements, single draws above),
    odd and even counts: synthetic keypoints related by a known homography + outliers"""
    rng = np.random.default_rng(m_count)
    kq = rng.integers(0, 3800, (m_count, 2)).astype(np.int32)
    Ht = np.array([[1.0, 0.01, 30.0], [-0.01, 1.0, 12.0], [1e-6, 2e-6, 1.0]])
    p = np.c_[kq, np.ones(m_count)] @ Ht.T
    kt = np.rint(p[:, :2] / p[:, 2:]).astype(np.int32)
    out = rng.random(m_count) < 0.4
    kt[out] = rng.integers(0, 3800, (int(out.sum()), 2))
    m = np.zeros(m_count, dtype=[("queryIdx", "<i4"), ("trainIdx", "<i4"), ("distance", "<f4")])
    m["queryIdx"] = m["trainIdx"] = np.arange(m_count)
    ro = oracle.ransac(kq, kt, m, iters=iters, seed=12345)
    Hr = ref.ransac(kq, kt, m, iters=iters, seed=12345)
    assert ro["ok"] and Hr is not None and np.array_equal(bits(ro["H"]), bits(Hr))


def test_reference_fold_matches_oracle_fold(ref, oracle, small):
    n, w, h, s, rs = (int(v) for v in small["fold_size"])
    views = load_synth().make_strip(n=n, w=w, h=h, seed=s)
    pano, log = oracle.stitch_fold(views, seed=rs)
    assert all(l["status"] == 1 for l in log)
    assert np.array_equal(pano, ref.stitch_all(views, seed=rs)["canvas"])


# ---------------- (2) oracle == committed reference outputs (runs anywhere) ------------------------
def test_oracle_equals_reference_golden_small(oracle, small):
    synth = load_synth()
    for tag in ("a", "b"):
        w, h, s = (int(v) for v in small["%s_size" % tag])
        left, right, _ = synth.make_pair(w, h, seed=s)
        kl, kr = oracle.detect(left), oracle.detect(right)
        assert np.array_equal(kl, small["%s_kl" % tag]) and np.array_equal(kr, small["%s_kr" % tag])
        m = oracle.match(kr, kl, right, left)
        assert np.array_equal(m["queryIdx"], small["%s_mq" % tag]) and np.array_equal(m["trainIdx"], small["%s_mt" % tag])
        assert np.array_equal(m["distance"], small["%s_ssd" % tag])
        r = oracle.ransac(kr, kl, m, seed=12345)
        assert np.array_equal(bits(r["H"]), bits(small["%s_H" % tag]))
        assert np.array_equal(oracle.stitch_pair(left, right, seed=12345)["canvas"], small["%s_canvas" % tag])
    n, w, h, s, rs = (int(v) for v in small["fold_size"])
    pano, _ = oracle.stitch_fold(synth.make_strip(n=n, w=w, h=h, seed=s), seed=rs)
    assert np.array_equal(pano, small["fold_canvas"])
    assert int(small["flat_status"]) == 0


def _check_run(o, left, right, g):
    kl, kr = o.detect(left), o.detect(right)
    assert (len(kl), len(kr)) == (g["kl"], g["kr"]) and sha(kl) == g["kl_sha"] and sha(kr) == g["kr_sha"]
    m = o.match(kr, kl, right, left)
    assert len(m) == g["m"] and sha(np.stack([m["queryIdx"], m["trainIdx"]], 1)) == g["m_sha"]
    assert sha(m["distance"]) == g["ssd_sha"]
    r = o.ransac(kr, kl, m, seed=g["seed"])
    assert hexbits(r["H"]) == g["H"]
    canvas = o.compose(left, right, r["H"])
    assert [canvas.shape[1], canvas.shape[0]] == g["canvas"] and sha(canvas) == g["canvas_sha"]


def test_oracle_equals_reference_golden_1080p(runs):
    from oracle.oracle import Oracle
    left, right, _ = load_synth().make_pair(1920, 1080, seed=31)
    _check_run(Oracle("omp"), left, right, runs["pair_1080p_seed31"])


def test_oracle_equals_reference_golden_c3_4k(runs, c3):
    """BASELINE config C3: the synthetic 3840x2160 pair, every stage, against the reference's digests"""
    o, left, right, kl, kr, m, r = c3
    g = runs["c3_pair_4k_seed267"]
    assert (len(kl), len(kr), len(m)) == (g["kl"], g["kr"], g["m"])
    assert sha(kl) == g["kl_sha"] and sha(kr) == g["kr_sha"] and sha(m["distance"]) == g["ssd_sha"]
    assert sha(np.stack([m["queryIdx"], m["trainIdx"]], 1)) == g["m_sha"]
    assert hexbits(r["H"]) == g["H"]
    canvas = o.compose(left, right, r["H"])
    assert [canvas.shape[1], canvas.shape[0]] == g["canvas"] and sha(canvas) == g["canvas_sha"]


def test_oracle_equals_reference_golden_c1_mountain(runs):
    cv2 = pytest.importorskip("cv2")
    from oracle.oracle import Oracle
    for base in (os.path.join(ROOT, "baseline", "_ref", "images"), "/root/reference/images"):
        p = [os.path.join(base, "mountain", "mountain%d.jpg" % i) for i in (1, 2)]
        if all(os.path.exists(q) for q in p):
            break
    else:
        pytest.skip("mountain sample images not present")
    if "c1_mountain" not in runs:
        pytest.skip("no reference run recorded for C1")
    _check_run(Oracle("omp"), cv2.imread(p[0]), cv2.imread(p[1]), runs["c1_mountain"])


# ---------------- (3) SURVEY c2-(ii): real cv2 at full size ------------------------------------------
@pytest.fixture(scope="module")
def c3():
    from oracle.oracle import Oracle
    o = Oracle("omp")
    left, right, _ = load_synth().make_pair(3840, 2160, seed=267)
    kl, kr = o.detect(left), o.detect(right)
    m = o.match(kr, kl, right, left)
    return o, left, right, kl, kr, m, o.ransac(kr, kl, m, seed=12345)


def test_full_size_gray_vs_cv2(c3):
    cv2 = pytest.importorskip("cv2")
    o, left, right = c3[:3]
    for img in (left, right):
        assert np.array_equal(o.gray(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))


def test_full_size_warp_and_overlay_vs_cv2(c3):
    """ref: src/serial/main.cpp:366-386 with the REAL cv2.warpPerspective on the 5763x2182 canvas"""
    cv2 = pytest.importorskip("cv2")
    o, left, right, kl, kr, m, r = c3
    ok, (cw, ch, ox, oy), TH = o.canvas_geometry(left.shape[1], left.shape[0], right.shape[1], right.shape[0], r["H"])
    assert ok
    pts = cv2.perspectiveTransform(np.float32([[[0, 0]], [[3840, 0]], [[3840, 2160]], [[0, 2160]]]), r["H"]).reshape(-1, 2)
    minx = min(np.float32(0), pts[:, 0].min()); miny = min(np.float32(0), pts[:, 1].min())
    T = np.array([[1, 0, -float(minx)], [0, 1, -float(miny)], [0, 0, 1.0]])
    assert np.array_equal(bits(cv2.gemm(T, r["H"], 1, None, 0)), bits(TH))
    warped = cv2.warpPerspective(right, TH, (cw, ch))
    canvas = np.zeros((ch, cw, 3), np.uint8)
    canvas[oy:oy + left.shape[0], ox:ox + left.shape[1]] = left
    nz = warped.any(axis=2)
    canvas[nz] = warped[nz]
    mine = o.compose(left, right, r["H"])
    # exact .5 ties in cv2's SIMD rounding are the only place the models may differ: assert none here
    assert np.array_equal(mine, canvas)


def test_all_ransac_hypotheses_vs_cv2_findhomography(c3):
    """all 1000 minimal samples of the C3 RANSAC run through the REAL cv2.findHomography: bit-equal H"""
    cv2 = pytest.importorskip("cv2")
    o, left, right, kl, kr, m, r = c3
    n_ok = 0
    for s in r["samples"]:
        src = kr[m["queryIdx"][s]].astype(np.float32)
        dst = kl[m["trainIdx"][s]].astype(np.float32)
        H, _ = cv2.findHomography(src, dst)
        Ho = o.find_homography4(src, dst)
        assert (H is None) == (Ho is None)
        if H is not None:
            assert np.array_equal(bits(H), bits(Ho))
            n_ok += 1
    assert n_ok > 900


def test_gemm_3x1_and_scaling_vs_cv2(c3):
    """H * (x, y, 1) (ref :288) and Mat /= w (ref :289) with the real cv2.gemm / convertScale... the inlier
    predicate's inputs: projections of every match under the best H, bit-equal"""
    cv2 = pytest.importorskip("cv2")
    o, left, right, kl, kr, m, r = c3
    H = r["H"]
    pts = kr[m["queryIdx"]][:2000].astype(np.float32)
    for x, y in pts[:300]:
        p = np.array([[float(x)], [float(y)], [1.0]])
        q = cv2.gemm(H, p, 1, None, 0)
        mine = np.array([(H[i, 0] * float(x) + H[i, 1] * float(y)) + H[i, 2] * 1.0 for i in range(3)])
        assert np.array_equal(bits(q.ravel()), bits(mine))
        s = cv2.multiply(q, 1.0, scale=1.0 / q[2, 0])       # Mat /= s  ==  convertTo(-1, 1./s)
        assert np.array_equal(bits(s.ravel()), bits(mine * (1.0 / mine[2])))
