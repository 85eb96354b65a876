"""GPU tier (-m gpu): the CUDA engine, through the C ABI, against outputs of the REFERENCE ITSELF.

tests/golden/ref_small.npz and ref_runs.json were produced by oracle/_ref — the unmodified
/root/reference/src/serial/main.cpp compiled against oracle/cvshim and run in the build container
(oracle/gen_ref_golden.py).  No oracle restatement sits between the two sides here: keypoints, match
lists, SSDs, homographies (bit-exact; bar 1e-4 relative) and canvases (identical; bar +-1 LSB) are
compared with what the reference's own code computed.  Where the prebuilt oracle/_ref travelled to the
GPU box, the reference is also run live next to the engine on fresh seeds.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_synth

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def hexbits(H):
    return [format(int(v), "016x") for v in bits(H).ravel()]


@pytest.fixture(scope="module")
def small():
    return np.load(os.path.join(GOLDEN, "ref_small.npz"))


@pytest.fixture(scope="module")
def runs():
    f = os.path.join(GOLDEN, "ref_runs.json")
    if not os.path.exists(f):
        pytest.skip("tests/golden/ref_runs.json not generated (oracle/gen_ref_golden.py)")
    return json.load(open(f))


def test_engine_equals_reference_small_vectors(engine, small):
    synth = load_synth()
    for tag in ("a", "b"):
        w, h, s = (int(v) for v in small["%s_size" % tag])
        left, right, _ = synth.make_pair(w, h, seed=s)
        kl = engine.gpuHarrisCornerDetectorDetect(left)
        kr = engine.gpuHarrisCornerDetectorDetect(right)
        assert np.array_equal(kl, small["%s_kl" % tag]) and np.array_equal(kr, small["%s_kr" % tag])
        m = engine.gpuHarrisMatchKeyPoints(kr, kl, right, left)
        assert np.array_equal(m["queryIdx"], small["%s_mq" % tag]) and np.array_equal(m["trainIdx"], small["%s_mt" % tag])
        assert np.array_equal(m["distance"], small["%s_ssd" % tag])
        H = engine.computeHomography(kr, kl, m)
        assert np.array_equal(bits(H), bits(small["%s_H" % tag]))
        canvas, r = engine.stitchTwoImages(left, right)
        assert r["status"] == 0 and np.array_equal(canvas, small["%s_canvas" % tag])
    n, w, h, s, rs = (int(v) for v in small["fold_size"])
    engine.set_seed(rs)
    try:
        pano, _ = engine.stitchAllImages(synth.make_strip(n=n, w=w, h=h, seed=s))
    finally:
        engine.set_seed(12345)
    assert np.array_equal(pano, small["fold_canvas"])


def _check(engine, left, right, g):
    kl = engine.gpuHarrisCornerDetectorDetect(left)
    kr = engine.gpuHarrisCornerDetectorDetect(right)
    assert (len(kl), len(kr)) == (g["kl"], g["kr"]) and sha(kl) == g["kl_sha"] and sha(kr) == g["kr_sha"]
    m = engine.gpuHarrisMatchKeyPoints(kr, kl, right, left)
    assert len(m) == g["m"] and sha(np.stack([m["queryIdx"], m["trainIdx"]], 1)) == g["m_sha"]
    assert sha(m["distance"]) == g["ssd_sha"]
    canvas, r = engine.stitchTwoImages(left, right)
    assert r["status"] == 0 and hexbits(r["H"]) == g["H"]
    assert [canvas.shape[1], canvas.shape[0]] == g["canvas"] and sha(canvas) == g["canvas_sha"]


def test_engine_equals_reference_1080p(engine, runs):
    left, right, _ = load_synth().make_pair(1920, 1080, seed=31)
    _check(engine, left, right, runs["pair_1080p_seed31"])


def test_engine_equals_reference_c3_4k_pair(engine, runs):
    """BASELINE config C3 (the benchmark pair): every stage against the reference's own run"""
    left, right, _ = load_synth().make_pair(3840, 2160, seed=267)
    _check(engine, left, right, runs["c3_pair_4k_seed267"])


def _photos(*names):
    cv2 = pytest.importorskip("cv2")
    paths = [os.path.join(ROOT, "baseline", "_ref", "images", n) for n in names]
    if not all(os.path.exists(p) for p in paths):
        pytest.skip("reference sample images not staged under baseline/_ref/images")
    return [cv2.imread(p) for p in paths]


def test_engine_equals_reference_c1_mountain(engine, runs):
    if "c1_mountain" not in runs:
        pytest.skip("no reference run recorded for C1")
    left, right = _photos("mountain/mountain1.jpg", "mountain/mountain2.jpg")
    _check(engine, left, right, runs["c1_mountain"])


def test_engine_equals_reference_c2_oilseed_fold(engine, runs):
    if "c2_oilseed_fold" not in runs:
        pytest.skip("no reference run recorded for C2")
    g = runs["c2_oilseed_fold"]
    ims = _photos(*["oilseed/oilseed%d.jpg" % i for i in g["order"]])
    engine.set_seed(g["seed"])
    try:
        pano, log = engine.stitchAllImages(ims)
    finally:
        engine.set_seed(12345)
    assert [pano.shape[1], pano.shape[0]] == g["canvas"] and sha(pano) == g["canvas_sha"]


@pytest.mark.parametrize("w,h,seed,rs", [(512, 300, 41, 3), (389, 277, 42, 12345), (800, 450, 43, 99)])
def test_engine_equals_reference_live(engine, w, h, seed, rs):
    """the prebuilt oracle/_ref (the reference's own code) run next to the engine on inputs that are in no fixture"""
    from oracle import ref as refmod
    if not refmod.available():
        pytest.skip("oracle/_ref did not travel to this box")
    R = refmod.Reference()
    left, right, _ = load_synth().make_pair(w, h, seed=seed)
    kl, kr = engine.gpuHarrisCornerDetectorDetect(left), engine.gpuHarrisCornerDetectorDetect(right)
    assert np.array_equal(kl, R.detect(left)) and np.array_equal(kr, R.detect(right))
    m = engine.gpuHarrisMatchKeyPoints(kr, kl, right, left)
    mr = R.match(kr, kl, right, left)
    assert np.array_equal(m["queryIdx"], mr["queryIdx"]) and np.array_equal(m["trainIdx"], mr["trainIdx"])
    assert np.array_equal(m["distance"], mr["distance"])
    engine.set_seed(rs)
    try:
        canvas, r = engine.stitchTwoImages(left, right)
    finally:
        engine.set_seed(12345)
    ref = R.stitch_pair(left, right, seed=rs)
    assert (r["status"] == 0) == (ref["status"] == 1)
    if ref["status"] == 1:
        assert np.array_equal(bits(r["H"]), bits(R.ransac(kr, kl, mr, seed=rs)))
        assert np.array_equal(canvas, ref["canvas"])
