"""GPU tier of the opt-in 2-NN / Lowe-ratio matcher (pano_match_knn; north star item (c); NOT a reference function,
never used by pano_stitch_*): the tensor-core matcher's top-2 epilogue (tcgen05, match_tc_top2_kernel), the SIMT
top-2 kernel and the binary-descriptor path (XOR / popcount, warp-shuffle reduction) against the checker
(oracle.match_knn, pinned to cv2.BFMatcher in tests/test_knn.py), through the C ABI, bit for bit.
(This file sorts last so that it runs after the parity suite of the reference's path.  These kernels were written
after the round's GPU budget was spent: on the CPU they run on the emulation tier of tests/test_knn.py; their first
run on a B200 is the driver's.)"""
import numpy as np
import pytest

from conftest import load_pkg, load_synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scenes(oracle):
    out = {}
    for name, (w, h, seed) in {"small": (640, 360, 11), "mid": (1920, 1080, 31)}.items():
        left, right, _ = load_synth().make_pair(w, h, seed=seed)
        kl, kr = oracle.detect(left), oracle.detect(right)
        # keypoints that fail the in-border test and duplicates (exact distance ties)
        kl = np.concatenate([kl, [[0, 0], [1, h - 1], [w - 1, 5]], kl[:9]]).astype(np.int32)
        kr = np.concatenate([[[2, 1]], kr, kr[3:6]]).astype(np.int32)
        out[name] = (left, right, kl, kr)
    return out


def check(engine, oracle, kq, kt, imq, imt, descriptor, ratio, patch=5, min_matches=0):
    m, s = engine.matchKnn(kq, kt, imq, imt, patchSize=patch, descriptor=descriptor, ratio=ratio)
    mo, so = oracle.match_knn(kq, kt, imq, imt, patch=patch, descriptor=descriptor, ratio=ratio)
    assert len(mo) >= min_matches
    assert len(m) == len(mo)
    assert np.array_equal(m["queryIdx"], mo["queryIdx"]) and np.array_equal(m["trainIdx"], mo["trainIdx"])
    assert np.array_equal(m["distance"], mo["distance"]) and np.array_equal(s, so)
    return m


SIMT, TC = 1, 0     # pano_set_matcher: 1 = SIMT top-2 kernel, 0 = tensor-core matcher with the top-2 epilogue (tcgen05)


class use_matcher:
    def __init__(self, engine, which):
        self.engine, self.which = engine, which

    def __enter__(self):
        self.engine.set_matcher(self.which)

    def __exit__(self, *a):
        self.engine.set_matcher(0)


# (ordered from the plain kernels to the tensor-core variant, so that a failure of the latter still leaves the
# results of the former in the log of a run with -x)
def test_knn_argument_checks(engine, scenes):
    pkg = load_pkg()
    left, right, kl, kr = scenes["small"]
    for kw, code in ((dict(ratio=0.0), pkg.PANO_ERR_INVALID), (dict(ratio=1.5), pkg.PANO_ERR_INVALID),
                     (dict(descriptor=7), pkg.PANO_ERR_UNSUPPORTED), (dict(descriptor=1, patchSize=3), pkg.PANO_ERR_UNSUPPORTED),
                     (dict(patchSize=4), pkg.PANO_ERR_UNSUPPORTED)):
        with pytest.raises(pkg.PanoError) as e:
            engine.matchKnn(kr, kl, right, left, **kw)
        assert e.value.status == code


def edge_cases(scenes):
    left, right, kl, kr = scenes["small"]
    w, h = 640, 360
    inb = ~((kl[:, 0] < 2) | (kl[:, 1] < 2) | (kl[:, 0] + 2 >= w) | (kl[:, 1] + 2 >= h))
    ok_l = kl[inb]
    return [(kr, ok_l[:1]),            # one candidate: no runner-up, no match
            (kr, ok_l[:2]), (kr[:1], kl), (kr[1:2], kl), (kr, kl[-12:]),
            (kr[:130], ok_l[:129]), (kr[:129], ok_l[:128]), (kr[:513], ok_l[:257]), (kr[:33], ok_l[:33])]


def flat_scene():
    flat = np.full((60, 80, 3), 77, np.uint8)      # every distance 0: the strict test rejects every query
    k = np.array([[x, y] for y in range(5, 50, 9) for x in range(5, 70, 7)], np.int32)
    return flat, k


def ssd_equals_checker(engine, oracle, scenes, matcher, scene):
    left, right, kl, kr = scenes[scene]
    with use_matcher(engine, matcher):
        for ratio in (0.75, 0.9, 1.0):
            check(engine, oracle, kr, kl, right, left, 0, ratio, min_matches=20)


def edge_cases_equal_checker(engine, oracle, scenes, descriptor, matcher):
    left, right, _, _ = scenes["small"]
    with use_matcher(engine, matcher):
        for kq, kt in edge_cases(scenes):
            check(engine, oracle, kq, kt, right, left, descriptor, 0.8)
        flat, k = flat_scene()
        m, _ = engine.matchKnn(k, k, flat, flat, descriptor=descriptor, ratio=1.0)
        assert len(m) == 0


def other_patch_size(engine, oracle, scenes, patch, matcher):
    left, right, kl, kr = scenes["small"]
    with use_matcher(engine, matcher):
        check(engine, oracle, kr, kl, right, left, 0, 0.9, patch=patch, min_matches=1)


# ---- SIMT top-2 kernel and the binary-descriptor path ---------------------------------------------------------------
@pytest.mark.parametrize("scene", ["small", "mid"])
def test_knn_simt_patch_ssd_equals_checker(engine, oracle, scenes, scene):
    ssd_equals_checker(engine, oracle, scenes, SIMT, scene)


@pytest.mark.parametrize("scene", ["small", "mid"])
def test_knn_binary_equals_checker(engine, oracle, scenes, scene):
    left, right, kl, kr = scenes[scene]
    for ratio in (0.8, 0.95, 1.0):
        check(engine, oracle, kr, kl, right, left, 1, ratio, min_matches=5)


@pytest.mark.parametrize("descriptor", [1, 0])
def test_knn_simt_edge_cases(engine, oracle, scenes, descriptor):
    edge_cases_equal_checker(engine, oracle, scenes, descriptor, SIMT)


@pytest.mark.parametrize("patch", [1, 3])
def test_knn_simt_other_patch_sizes(engine, oracle, scenes, patch):
    other_patch_size(engine, oracle, scenes, patch, SIMT)


# ---- the tensor-core matcher's top-2 epilogue ----------------------------------------------------------------------
@pytest.mark.parametrize("scene", ["small", "mid"])
def test_knn_tc_patch_ssd_equals_checker(engine, oracle, scenes, scene):
    ssd_equals_checker(engine, oracle, scenes, TC, scene)


def test_knn_tc_edge_cases(engine, oracle, scenes):
    edge_cases_equal_checker(engine, oracle, scenes, 0, TC)


@pytest.mark.parametrize("patch", [1, 3])
def test_knn_tc_other_patch_sizes(engine, oracle, scenes, patch):
    other_patch_size(engine, oracle, scenes, patch, TC)


def test_knn_tensor_core_and_simt_agree_with_the_reference_matcher_at_ratio_one(engine, scenes):
    """with ratio 1 every query whose nearest neighbour is strictly nearer than the runner-up passes: those rows are
    the rows of the reference's matcher (pano_match; ref: src/serial/main.cpp:188-244)"""
    left, right, kl, kr = scenes["mid"]
    ref = {int(r["queryIdx"]): r for r in engine.gpuHarrisMatchKeyPoints(kr, kl, right, left)}
    for matcher in (SIMT, TC):
        with use_matcher(engine, matcher):
            m, s = engine.matchKnn(kr, kl, right, left, ratio=1.0)
        assert 0 < len(m) <= len(ref)
        for rec, sec in zip(m, s):
            r = ref[int(rec["queryIdx"])]
            assert (rec["trainIdx"], rec["distance"]) == (r["trainIdx"], r["distance"]) and rec["distance"] < sec


def test_knn_4k_tensor_core_top2(engine, oracle):
    """BASELINE config 3's size: ~11 k x 11 k candidates on the tensor cores, every (nearest, runner-up) pair exact"""
    left, right, _ = load_synth().make_pair(3840, 2160, seed=267)
    kl = engine.gpuHarrisCornerDetectorDetect(left)
    kr = engine.gpuHarrisCornerDetectorDetect(right)
    assert len(kl) > 5000 and len(kr) > 5000
    m = check(engine, oracle, kr, kl, right, left, 0, 0.75, min_matches=100)
    # the stitching path is untouched by the opt-in call: the reference's matcher still returns one row per query
    ref = engine.gpuHarrisMatchKeyPoints(kr, kl, right, left)
    assert len(ref) >= len(m) and set(m["queryIdx"]) <= set(ref["queryIdx"])
