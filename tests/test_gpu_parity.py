"""GPU tier (-m gpu): the CUDA engine, called through the C ABI, against the CPU oracle on the
same inputs.  Bars (BASELINE.json north_star): keypoints, match lists, RANSAC samples / counts /
inlier sets bit-exact; homographies within 1e-4 relative (asserted bit-exact here, which is
stronger); warped pixels within +-1 LSB (asserted exact)."""
import os
import numpy as np
import pytest

from conftest import load_pkg, load_synth

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


# ---------------- K1 / K2: detector --------------------------------------------------------
def test_harris_response_bit_exact(engine, oracle, small_pair):
    left, right, _ = small_pair
    for img in (left, right):
        r_gpu = engine.harrisResponse(img)
        r_cpu = oracle.harris_response(img)
        assert np.array_equal(bits(r_gpu), bits(r_cpu))


@pytest.mark.parametrize("w,h", [(64, 48), (33, 35), (97, 61), (130, 5), (5, 130), (6, 6), (255, 257)])
def test_detect_ragged_sizes(engine, oracle, w, h):
    rng = np.random.default_rng(w * 1000 + h)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    # blocky content so that corners exist
    img[h // 4: h // 2, w // 4: w // 2] = (250, 20, 30)
    assert np.array_equal(bits(engine.harrisResponse(img)), bits(oracle.harris_response(img)))
    assert np.array_equal(engine.gpuHarrisCornerDetectorDetect(img), oracle.detect(img))


def test_detect_keypoints_identical_and_ordered(engine, oracle, small_pair, mid_pair):
    for pair in (small_pair, mid_pair):
        for img in pair[:2]:
            k_gpu = engine.gpuHarrisCornerDetectorDetect(img)
            k_cpu = oracle.detect(img)
            assert len(k_cpu) > 100
            assert np.array_equal(k_gpu, k_cpu)


def test_detect_flat_image_has_no_keypoints(engine, oracle):
    img = np.full((70, 90, 3), 128, np.uint8)
    assert len(engine.gpuHarrisCornerDetectorDetect(img)) == 0 == len(oracle.detect(img))


def test_detect_ties_are_rejected(engine, oracle):
    """strict NMS (ref: src/serial/main.cpp:164-176): periodic texture creates exact ties"""
    img = np.zeros((96, 96, 3), np.uint8)
    img[::8, :, :] = 255
    img[:, ::8, :] = 255
    assert np.array_equal(engine.gpuHarrisCornerDetectorDetect(img), oracle.detect(img))


def test_detect_other_options(engine, oracle, small_pair):
    left = small_pair[0]
    for k, th, nb in ((0.06, 5e5, 3), (0.04, 1e7, 5), (0.04, 1e5, 7)):
        assert np.array_equal(engine.gpuHarrisCornerDetectorDetect(left, k, th, nb), oracle.detect(left, k, th, nb))


def test_convolve_matches_oracle(engine, oracle):
    rng = np.random.default_rng(1)
    a = rng.uniform(-1000, 1000, (45, 67))
    for ks in (3, 5, 7):
        kern = rng.uniform(-1, 1, (ks, ks))
        assert np.array_equal(bits(engine.convolveCUDA(a, kern)), bits(oracle.convolve(a, kern)))


# ---------------- K3 / K4: matcher -----------------------------------------------------------
@pytest.mark.parametrize("which", [1, 0])
def test_matches_identical(engine, oracle, small_pair, mid_pair, which):
    engine.set_matcher(which)
    try:
        for left, right, _ in (small_pair, mid_pair):
            kl, kr = oracle.detect(left), oracle.detect(right)
            m_cpu = oracle.match(kr, kl, right, left)
            m_gpu = engine.gpuHarrisMatchKeyPoints(kr, kl, right, left)
            assert len(m_cpu) > 50
            assert np.array_equal(m_gpu, m_cpu)
    finally:
        engine.set_matcher(0)


@pytest.mark.parametrize("which", [1, 0])
def test_match_edge_cases(engine, oracle, small_pair, which):
    engine.set_matcher(which)
    try:
        left, right, _ = small_pair
        kl, kr = oracle.detect(left), oracle.detect(right)
        # keypoints on the border are skipped on both sides; duplicates create exact SSD ties
        kq = np.concatenate([[[0, 0], [1, 5], [right.shape[1] - 1, 9], [5, right.shape[0] - 2]], kr[:300]]).astype(np.int32)
        kt = np.concatenate([kl[:200], kl[:200], [[1, 1], [left.shape[1] - 2, 7]]]).astype(np.int32)
        for off in (0, 17):
            a = engine.gpuHarrisMatchKeyPoints(kq, kt, right, left, offset=off)
            b = oracle.match(kq, kt, right, left, offset=off)
            assert np.array_equal(a, b)
        # a threshold that really filters
        a = engine.gpuHarrisMatchKeyPoints(kr, kl, right, left, maxSSDThresh=3000.0)
        b = oracle.match(kr, kl, right, left, max_ssd=3000.0)
        assert 0 < len(b) < len(kr) and np.array_equal(a, b)
        # other patch size
        a = engine.gpuHarrisMatchKeyPoints(kr, kl, right, left, patchSize=3)
        b = oracle.match(kr, kl, right, left, patch=3)
        assert np.array_equal(a, b)
        # empty sides
        empty = np.zeros((0, 2), np.int32)
        assert len(engine.gpuHarrisMatchKeyPoints(empty, kl, right, left)) == 0
        assert len(engine.gpuHarrisMatchKeyPoints(kr, empty, right, left)) == 0
        # no in-border train keypoint at all
        assert len(engine.gpuHarrisMatchKeyPoints(kr, np.array([[0, 0], [1, 1]], np.int32), right, left)) == 0
    finally:
        engine.set_matcher(0)


# ---------------- K5 - K7: RANSAC ------------------------------------------------------------
def _ransac_inputs(oracle, pair):
    left, right, _ = pair
    kl, kr = oracle.detect(left), oracle.detect(right)
    return kl, kr, oracle.match(kr, kl, right, left)


@pytest.mark.parametrize("seed", [12345, 1, 267])
def test_ransac_bit_exact(engine, oracle, small_pair, seed):
    kl, kr, m = _ransac_inputs(oracle, small_pair)
    engine.set_seed(seed)
    g = engine.computeHomography(kr, kl, m, details=True)
    c = oracle.ransac(kr, kl, m, seed=seed)
    engine.set_seed(12345)
    assert g["ok"] and c["ok"]
    assert np.array_equal(g["samples"], c["samples"])          # replayed std::shuffle
    assert np.array_equal(g["counts"], c["counts"])            # per-iteration inlier counts
    assert (g["best_count"], g["best_iter"]) == (c["best_count"], c["best_iter"])
    assert np.array_equal(g["inlier_mask"], c["inlier_mask"])  # inlier set
    assert np.array_equal(bits(g["H"]), bits(c["H"]))          # homography (bar: 1e-4 relative)


def test_ransac_mid_size_and_subsets(engine, oracle, mid_pair):
    kl, kr, m = _ransac_inputs(oracle, mid_pair)
    pkg = load_pkg()
    for sub in (len(m), 1001, 4, 5, 6, 7, 8, 64):   # even / odd / tiny match counts
        mm = m[:sub]
        g = engine.computeHomography(kr, kl, mm, details=True)
        c = oracle.ransac(kr, kl, mm, seed=12345)
        assert g["ok"] == c["ok"]
        assert np.array_equal(g["samples"], c["samples"])
        assert np.array_equal(g["counts"], c["counts"])
        if c["ok"]:
            assert np.array_equal(bits(g["H"]), bits(c["H"]))
            assert np.array_equal(g["inlier_mask"], c["inlier_mask"])
    # fewer matches than samples: the reference's loop breaks at once -> empty H
    assert engine.computeHomography(kr, kl, m[:3]) is None
    # fewer iterations
    o = pkg.RansacOptions(numIterations_=37)
    g = engine.computeHomography(kr, kl, m, o, details=True)
    c = oracle.ransac(kr, kl, m, iters=37, seed=12345)
    assert np.array_equal(g["samples"], c["samples"][:37]) and np.array_equal(g["counts"], c["counts"][:37])


def test_ransac_degenerate_samples(engine, oracle):
    """all matches share an x (findHomography returns empty every time) -> no homography"""
    kp1 = np.array([[5, i * 3] for i in range(40)], np.int32)
    kp2 = np.array([[9 + i, i * 3 + 1] for i in range(40)], np.int32)
    m = np.zeros(40, load_pkg().MATCH_DTYPE)
    m["queryIdx"] = np.arange(40); m["trainIdx"] = np.arange(40)
    g = engine.computeHomography(kp1, kp2, m, details=True)
    c = oracle.ransac(kp1, kp2, m, seed=12345)
    assert not g["ok"] and not c["ok"]
    assert np.array_equal(g["counts"], c["counts"]) and (c["counts"] == -1).all()


# ---------------- K8: warp + overlay ---------------------------------------------------------
def test_warp_perspective_matches_cv2_fixtures(engine, pins):
    for i in range(4):
        ref = pins["warp_out%d" % i]
        out = engine.warpPerspective(pins["warp_src"], pins["warp_M%d" % i], (ref.shape[1], ref.shape[0]))
        assert np.array_equal(out, ref)


def test_warp_overlay_identical(engine, oracle, small_pair):
    left, right, Ht = small_pair
    rng = np.random.default_rng(4)
    Hs = [Ht, np.array([[1.0, 0, 480.0], [0, 1, 0], [0, 0, 1]]), np.array([[1.0, 0, -100.5], [0, 1, -20.25], [0, 0, 1]]),
          Ht @ np.array([[1.01, 0.01, 3], [-0.01, 0.99, -40], [1e-6, 2e-6, 1]])]
    for H in Hs:
        a = engine.warpOverlay(left, right, H)
        b = oracle.compose(left, right, H)
        assert (a is None) == (b is None)
        if b is not None:
            assert a.shape == b.shape
            assert np.abs(a.astype(int) - b.astype(int)).max() <= 1   # the bar
            assert np.array_equal(a, b)                               # what we actually get


# ---------------- fused pair / fold ------------------------------------------------------------
def test_stitch_pair_end_to_end(engine, oracle, small_pair, mid_pair):
    for left, right, Ht in (small_pair, mid_pair):
        canvas, r = engine.stitchTwoImages(left, right)
        o = oracle.stitch_pair(left, right, seed=12345)
        assert r["status"] == 0 and o["status"] == 1
        assert (r["kl"], r["kr"], r["m"], r["best"]) == (o["stats"]["kl"], o["stats"]["kr"], o["stats"]["m"], o["stats"]["best"])
        assert np.array_equal(bits(r["H"]), bits(o["H"]))
        assert np.abs(r["H"] / r["H"][2, 2] - Ht).max() / np.abs(Ht).max() < 5e-3     # recovers the truth
        assert r["canvas"] == o["geom"]
        assert np.array_equal(canvas, o["canvas"])


def test_stitch_pair_device_resident_inputs(engine, oracle, small_pair):
    import torch
    left, right, _ = small_pair
    L, R = torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()
    canvas, r = engine.stitchTwoImages(L, R)
    o = oracle.stitch_pair(left, right, seed=12345)
    assert r["status"] == 0
    assert np.array_equal(canvas.cpu().numpy(), o["canvas"])


def test_stitch_failure_statuses(engine, oracle):
    pkg = load_pkg()
    flat = np.full((64, 80, 3), 90, np.uint8)
    canvas, r = engine.stitchTwoImages(flat, flat)
    assert canvas is None and r["status"] == pkg.PANO_ERR_NO_MATCHES
    assert oracle.stitch_pair(flat, flat)["status"] == 0


def test_stitch_fold_three_images(engine, oracle):
    views = load_synth().make_strip(n=3, w=640, h=400, seed=5)
    pano, log = engine.stitchAllImages(views)
    opano, olog = oracle.stitch_fold(views, seed=12345)
    assert [l["status"] == 0 for l in log] == [l["status"] == 1 for l in olog]
    assert pano.shape == opano.shape and np.array_equal(pano, opano)


# ---------------- full-size properties (BASELINE config 3: 3840x2160 pair) ----------------------
def test_full_size_pair_properties(engine, oracle):
    left, right, Ht = load_synth().make_pair(3840, 2160, seed=267)
    canvas, r = engine.stitchTwoImages(left, right)
    assert r["status"] == 0
    H = r["H"] / r["H"][2, 2]
    assert np.abs(H - Ht).max() / np.abs(Ht).max() < 2e-3
    assert 8000 < r["kl"] < 25000 and r["m"] > 4000 and r["best"] > 0.3 * r["m"]
    # idempotence / determinism: same inputs, same seed -> identical outputs
    canvas2, r2 = engine.stitchTwoImages(left, right)
    assert np.array_equal(canvas, canvas2) and np.array_equal(bits(r["H"]), bits(r2["H"]))
    # the left image appears verbatim wherever the warped right image is black
    cw, ch, ox, oy = r["canvas"]
    region = canvas[oy:oy + left.shape[0], ox:ox + 200]   # far left: right image does not reach
    assert np.array_equal(region, left[:, :200])
    # stage parity at full size against the oracle (keypoints, matches, RANSAC)
    kl, kr = oracle.detect(left), oracle.detect(right)
    assert np.array_equal(engine.gpuHarrisCornerDetectorDetect(left), kl)
    m = oracle.match(kr, kl, right, left)
    assert np.array_equal(engine.gpuHarrisMatchKeyPoints(kr, kl, right, left), m)
    c = oracle.ransac(kr, kl, m, seed=12345)
    assert np.array_equal(bits(r["H"]), bits(c["H"])) and r["best"] == c["best_count"]
    assert np.array_equal(canvas, oracle.compose(left, right, c["H"]))


# ---------------- chain mode ---------------------------------------------------
def test_chain_of_two_equals_pair(engine, oracle, small_pair):
    left, right, _ = small_pair
    pano, res = engine.stitchChain([left, right])
    canvas, r = engine.stitchTwoImages(left, right)
    assert res[0]["status"] == 0 and np.array_equal(bits(res[0]["H"]), bits(r["H"]))
    assert np.array_equal(pano, canvas)


def test_chain_strip_matches_oracle_and_bands_tile(engine, oracle):
    views = load_synth().make_strip(n=4, w=640, h=400, seed=9)
    pano, res = engine.stitchChain(views)
    opano, oH = oracle.stitch_chain(views, seed=12345)
    assert [r["status"] == 0 for r in res] == [h is not None for h in oH]
    for r, h in zip(res, oH):
        if h is not None:
            assert np.array_equal(bits(r["H"]), bits(h))
    assert pano.shape == opano.shape and np.array_equal(pano, opano)
    # canvas rows rendered band by band (what each GPU does in the distributed mode) tile exactly
    Hs = engine.composeChain([r["H"] if r["status"] == 0 else None for r in res])
    ok, geom, T = engine.chainGeometry([(v.shape[1], v.shape[0]) for v in views], Hs)
    assert ok and (geom[0], geom[1]) == (pano.shape[1], pano.shape[0])
    bands, y = [], 0
    for bh in (7, 100, 1, geom[1] - 108):
        bands.append(engine.renderChainBand(views, Hs, geom, T, y, bh))
        y += bh
    assert np.array_equal(np.concatenate(bands, axis=0), pano)


# ---------------- replay regimes at large match counts (synthetic matches, no images needed) -------
@pytest.mark.parametrize("m,iters", [(40001, 60), (65535, 24), (65536, 24), (70000, 30)])
def test_ransac_large_match_counts(engine, oracle, m, iters):
    """paired draws with heavy Lemire rejection (ranges near 2^32) and, above 65535 elements,
    libstdc++'s one-draw-per-element branch; samples and counts must still be bit-exact"""
    rng = np.random.default_rng(m)
    kp1 = rng.integers(0, 4000, (m, 2)).astype(np.int32)
    kp2 = (kp1 + np.array([900, 3]) + rng.integers(-40, 41, (m, 2))).astype(np.int32)
    good = rng.random(m) < 0.5
    kp2[good] = kp1[good] + np.array([900, 3])
    mt = np.zeros(m, load_pkg().MATCH_DTYPE)
    mt["queryIdx"] = np.arange(m); mt["trainIdx"] = np.arange(m)
    o = load_pkg().RansacOptions(numIterations_=iters)
    g = engine.computeHomography(kp1, kp2, mt, o, details=True)
    c = oracle.ransac(kp1, kp2, mt, iters=iters, seed=12345)
    assert np.array_equal(g["samples"], c["samples"][:iters])
    assert np.array_equal(g["counts"], c["counts"][:iters])
    assert g["ok"] == c["ok"] and (not c["ok"] or np.array_equal(bits(g["H"]), bits(c["H"])))


def test_replay_window_miss_is_detected_and_rerun(oracle, small_pair):
    """with absurdly narrow speculation windows (0.3 sigma) the true start falls outside its window
    all the time; the engine must notice, widen and still return the exact samples"""
    import subprocess
    import sys
    import textwrap
    import os
    from conftest import ROOT, PKG
    code = textwrap.dedent('''
        import sys, importlib, numpy as np
        sys.path.insert(0, %r)
        pkg = importlib.import_module(%r); synth = importlib.import_module(%r + ".synth")
        from oracle.oracle import Oracle
        O = Oracle(); eng = pkg.Engine(0, 12345)
        left, right, _ = synth.make_pair(960, 540, seed=267)
        kl, kr = O.detect(left), O.detect(right); m = O.match(kr, kl, right, left)
        g = eng.computeHomography(kr, kl, m, details=True); c = O.ransac(kr, kl, m, seed=12345)
        assert np.array_equal(g["samples"], c["samples"]) and np.array_equal(g["counts"], c["counts"])
        # the fused pair call pre-launches the replay on a side stream: a miss there must fall back to the in-order
        # re-run as well
        canvas, r = eng.stitchTwoImages(left, right); o = O.stitch_pair(left, right, seed=12345)
        assert r["status"] == 0 and np.array_equal(canvas, o["canvas"])
        assert np.array_equal(r["H"].view(np.uint64), o["H"].view(np.uint64))
        print("OK", len(m))
    ''') % (ROOT, PKG, PKG)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                       env=dict(os.environ, PANO_REPLAY_Z="0.3"))
    assert p.returncode == 0 and "OK" in p.stdout, p.stderr[-800:]


# ---------------- resident replay (one CTA, iterations in order; pano_stitch_batch's lanes) --------
@pytest.mark.parametrize("m,iters", [(4, 50), (5, 50), (9, 100), (257, 300), (3000, 1000), (10774, 1000),
                                     (10837, 1000), (20001, 100), (40001, 30)])
def test_ransac_resident_replay(engine, oracle, m, iters):
    """pano_set_replay_mode(1): samples, counts and H identical to the oracle (and so to the chunked
    replay); m = 20001 and 40001 exceed the resident plan's limits and must fall back to the chunked path"""
    rng = np.random.default_rng(m)
    kp1 = rng.integers(0, 4000, (m, 2)).astype(np.int32)
    kp2 = (kp1 + np.array([900, 3]) + rng.integers(-40, 41, (m, 2))).astype(np.int32)
    good = rng.random(m) < 0.5
    kp2[good] = kp1[good] + np.array([900, 3])
    mt = np.zeros(m, load_pkg().MATCH_DTYPE)
    mt["queryIdx"] = np.arange(m); mt["trainIdx"] = np.arange(m)
    o = load_pkg().RansacOptions(numIterations_=iters)
    engine.set_replay_mode(1)
    try:
        g = engine.computeHomography(kp1, kp2, mt, o, details=True)
    finally:
        engine.set_replay_mode(0)
    c = oracle.ransac(kp1, kp2, mt, iters=iters, seed=12345)
    assert np.array_equal(g["samples"], c["samples"][:iters])
    assert np.array_equal(g["counts"], c["counts"][:iters])
    assert g["ok"] == c["ok"] and (not c["ok"] or np.array_equal(bits(g["H"]), bits(c["H"])))


def test_resident_band_miss_is_detected_and_rerun(oracle):
    """narrow bands (PANO_REPLAY_Z=0.3 -> 0.6 sigma) make paths leave their band; the kernel must report
    it and the engine re-plan wider, never return guessed samples"""
    import subprocess
    import sys
    import textwrap
    import os
    from conftest import ROOT, PKG
    code = textwrap.dedent('''
        import sys, importlib, numpy as np
        sys.path.insert(0, %r)
        pkg = importlib.import_module(%r); synth = importlib.import_module(%r + ".synth")
        from oracle.oracle import Oracle
        O = Oracle(); eng = pkg.Engine(0, 12345); eng.set_replay_mode(1)
        left, right, _ = synth.make_pair(960, 540, seed=267)
        kl, kr = O.detect(left), O.detect(right); m = O.match(kr, kl, right, left)
        g = eng.computeHomography(kr, kl, m, details=True); c = O.ransac(kr, kl, m, seed=12345)
        assert np.array_equal(g["samples"], c["samples"]) and np.array_equal(g["counts"], c["counts"])
        print("OK", len(m))
    ''') % (ROOT, PKG, PKG)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                       env=dict(os.environ, PANO_REPLAY_Z="0.3"))
    assert p.returncode == 0 and "OK" in p.stdout, p.stderr[-800:]


# ---------------- throughput mode: pano_stitch_batch (lanes, prefetched uploads, async downloads) ----
@pytest.mark.parametrize("mem", ["host", "device"])
def test_stitch_batch_equals_pairs(engine, oracle, mem):
    """every pair of a batch (several pairs per lane, host buffers with upload / download streams, or device
    buffers) gives exactly what pano_stitch_pair gives for it: H, counts and canvas bytes"""
    import torch
    synth = load_synth()
    pairs = [synth.make_pair(640, 360, seed=40 + i)[:2] for i in range(7)]
    lefts, rights = [p[0] for p in pairs], [p[1] for p in pairs]
    cap = 3 * (2 * 640 + 64) * (360 + 256)
    if mem == "host":
        L, R = lefts, rights
        outs = [np.zeros(cap, np.uint8) for _ in pairs]
    else:
        L = [torch.from_numpy(a).cuda() for a in lefts]
        R = [torch.from_numpy(a).cuda() for a in rights]
        outs = [torch.zeros(cap, dtype=torch.uint8, device="cuda") for _ in pairs]
    os.environ["PANO_BATCH_LANES"] = "3"      # 7 pairs on 3 lanes: up to 3 pairs per lane (slot reuse, canvas reuse)
    try:
        res, _ = engine.stitchBatch(L, R, canvases_out=outs)
    finally:
        os.environ.pop("PANO_BATCH_LANES", None)
    for i, r in enumerate(res):
        canvas, p = engine.stitchTwoImages(lefts[i], rights[i])
        assert r["status"] == p["status"] == 0
        assert np.array_equal(bits(r["H"]), bits(p["H"])) and r["best"] == p["best"] and r["m"] == p["m"]
        cw, ch = r["canvas"][0], r["canvas"][1]
        got = outs[i] if mem == "host" else outs[i].cpu().numpy()
        assert np.array_equal(got[:cw * ch * 3].reshape(ch, cw, 3), canvas)


def test_general_warp_kernel_equals_fast_path(oracle):
    """PANO_WARP_FAST=0 forces the general (exact-division) kernel, PANO_WARP_KERNEL=1 the round-1 one-pixel-per-lane
    fast kernel, the default is the quad kernel (4 px per thread); all three must reproduce the oracle's canvas (pair
    overlay, plain warpPerspective, band-wise accumulate of chain mode), on odd canvas widths and unaligned strides"""
    import subprocess
    import sys
    import textwrap
    from conftest import ROOT, PKG
    code = textwrap.dedent('''
        import sys, importlib, numpy as np
        sys.path.insert(0, %r)
        pkg = importlib.import_module(%r); synth = importlib.import_module(%r + ".synth")
        from oracle.oracle import Oracle
        O = Oracle(); eng = pkg.Engine(0, 12345)
        left, right, _ = synth.make_pair(960, 540, seed=5)
        canvas, r = eng.stitchTwoImages(left, right)
        o = O.stitch_pair(left, right, seed=12345)
        assert r["status"] == 0 and np.array_equal(canvas, o["canvas"])
        M = np.array([[0.98, 0.03, 40.5], [-0.02, 1.01, 12.25], [1e-5, -2e-5, 1.0]])
        for ds in ((1100, 640), (1101, 77), (333, 9), (130, 641)):
            w = eng.warpPerspective(right, M, ds)
            assert np.array_equal(w, O.warp_perspective(right, M, ds)), ds
        views = synth.make_strip(n=3, w=500, h=300, seed=4)
        c1, _ = eng.stitchChain(views)
        c2, _ = O.stitch_chain(views, seed=12345)
        assert np.array_equal(c1, c2)
        l2, r2, _ = synth.make_pair(641, 363, seed=8)           # odd width: 3 * w is not a multiple of 4
        cv, rr = eng.stitchTwoImages(l2, r2)
        oo = O.stitch_pair(l2, r2, seed=12345)
        assert (rr["status"] == 0) == (oo["status"] == 1) and (oo["status"] != 1 or np.array_equal(cv, oo["canvas"]))
        print("OK")
    ''') % (ROOT, PKG, PKG)
    for fast, kern in (("0", "0"), ("1", "0"), ("1", "1")):
        p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                           env=dict(os.environ, PANO_WARP_FAST=fast, PANO_WARP_KERNEL=kern))
        assert p.returncode == 0 and "OK" in p.stdout, (fast, kern, p.stderr[-800:])


def test_ransac_rejects_out_of_range_match_indices(engine, oracle, small_pair):
    """pano_ransac checks the caller's match indices against n1 / n2 on the device: an index outside the keypoint
    lists (the reference would read past its vectors) is reported as PANO_ERR_INVALID, never dereferenced, and the
    context keeps working afterwards"""
    pkg = load_pkg()
    left, right, _ = small_pair
    kl, kr = oracle.detect(left), oracle.detect(right)
    m = oracle.match(kr, kl, right, left)
    for field, bad in (("queryIdx", len(kr) + 3), ("trainIdx", len(kl)), ("trainIdx", -1)):
        mb = m.copy()
        mb[field][7] = bad
        with pytest.raises(pkg.PanoError) as ei:
            engine.computeHomography(kr, kl, mb)
        assert ei.value.status == pkg.PANO_ERR_INVALID
    H = engine.computeHomography(kr, kl, m)
    assert np.array_equal(bits(H), bits(oracle.ransac(kr, kl, m, seed=12345)["H"]))


def test_incremental_fold_matches_its_restatement(oracle):
    """opt-in fold restructuring (pano_set_fold_mode(ctx, 1)): only the new image is detected, the
    panorama's keypoints are carried (shifted / mapped through T*H).  Behaviour differs from the reference's fold by
    design, so the engine is held to a restatement of the same rule built from the oracle's stage functions."""
    pkg, synth = load_pkg(), load_synth()
    views = synth.make_strip(n=4, w=520, h=300, seed=12)
    eng = pkg.Engine(0, 12345)
    eng.set_fold_mode(1)
    pano, log = eng.stitchAllImages(views)
    eng.set_fold_mode(0)
    ref_pano, _ = eng.stitchAllImages(views)
    eng.close()
    O = oracle
    P, K = views[0], O.detect(views[0])
    for im, step in zip(views[1:], log):
        kr = O.detect(im)
        m = O.match(kr, K, im, P)
        r = O.ransac(kr, K, m, seed=12345)
        assert (step["kl"], step["kr"], step["m"], step["best"]) == (len(K), len(kr), len(m), r["best_count"])
        assert r["ok"] and step["status"] == 0 and np.array_equal(bits(step["H"]), bits(r["H"]))
        ok, (cw, ch, ox, oy), TH = O.canvas_geometry(P.shape[1], P.shape[0], im.shape[1], im.shape[0], r["H"])
        assert ok
        P = O.compose(P, im, r["H"])
        moved = np.where(K[:, :1] < 0, -1, K + np.int32([ox, oy])).astype(np.int32)
        q = np.rint(O.perspective_transform(kr.astype(np.float32), TH)).astype(np.int32)
        inside = (q[:, 0] >= 0) & (q[:, 1] >= 0) & (q[:, 0] < cw) & (q[:, 1] < ch)
        q[~inside] = -1
        K = np.concatenate([moved, q]).astype(np.int32)
    assert pano.shape == P.shape and np.array_equal(pano, P)
    # it is a different algorithm from the reference's fold (which re-detects on the panorama), same kind of result
    assert abs(pano.shape[1] - ref_pano.shape[1]) < 0.15 * ref_pano.shape[1], (pano.shape, ref_pano.shape)


def test_async_pair_equals_blocking_pair(engine, oracle, small_pair):
    """pano_stitch_pair_async + pano_pair_query / pano_pair_wait: same result as the blocking call; the
    context refuses other work while the pair is in flight"""
    pkg = load_pkg()
    import torch
    left, right, _ = small_pair
    o = oracle.stitch_pair(left, right, seed=12345)
    h = engine.stitchTwoImagesAsync(left, right)
    res2 = pkg.PairResult()
    ho, ro = pkg.HarrisCornerOptions(), pkg.RansacOptions()
    L = pkg._Img(left)
    import ctypes as C
    busy = engine.lib.pano_stitch_pair_async(engine.ctx, L.ptr, L.w, L.h, C.c_size_t(L.stride), L.ptr, L.w, L.h,
                                             C.c_size_t(L.stride), 0, C.byref(ho), C.byref(ro), None, C.byref(res2))
    assert busy in (pkg.PANO_ERR_BUSY,)
    r = h.result()
    assert h.done() and r["status"] == 0 and np.array_equal(bits(r["H"]), bits(o["H"]))
    assert np.array_equal(engine.getCanvas(), o["canvas"])
    # stream-ordered, device-resident: work enqueued on the caller's stream after the wait sees the canvas
    Ld, Rd = torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()
    s = torch.cuda.Stream()
    h = engine.stitchTwoImagesAsync(Ld, Rd, stream_ptr=s.cuda_stream)
    r = h.result()
    with torch.cuda.stream(s):
        c = engine.getCanvas(device=True)
    s.synchronize()
    assert r["status"] == 0 and np.array_equal(c.cpu().numpy(), o["canvas"])
