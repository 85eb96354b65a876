"""GPU tier: the opt-in matcher of the fused calls (pano_set_match_mode(ctx, 1, ratio, descriptor): 2 nearest neighbours
+ Lowe's ratio test feed RANSAC instead of the reference's nearest-patch matches; off by default, behaviour changing by
design).  The fused pair, the fold and the batch must equal the composition of the checker's stage functions (detect,
match_knn, ransac, compose), and mode 0 must be the reference again.  The same checks run on the emulated engine in the
CPU tier (tests/test_engine_emu.py); written after the round's GPU budget was spent - first run on a B200: the driver's."""
import numpy as np
import pytest

from conftest import load_pkg, load_synth

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def oracle_pair_knn(oracle, left, right, ratio, descriptor, iters=1000, seed=12345):
    kl, kr = oracle.detect(left), oracle.detect(right)
    m, _ = oracle.match_knn(kr, kl, right, left, descriptor=descriptor, ratio=ratio)
    o = oracle.ransac(kr, kl, m, iters=iters, seed=seed)
    return o, len(kl), len(kr), len(m), (oracle.compose(left, right, o["H"]) if o["ok"] else None)


@pytest.mark.parametrize("descriptor,ratio", [(0, 0.75), (1, 0.9)])
def test_pair_with_the_ratio_test_matcher(engine, oracle, small_pair, descriptor, ratio):
    left, right, _ = small_pair
    engine.set_match_mode(1, ratio=ratio, descriptor=descriptor)
    try:
        canvas, r = engine.stitchTwoImages(left, right)
    finally:
        engine.set_match_mode(0)
    o, nkl, nkr, nm, want = oracle_pair_knn(oracle, left, right, ratio, descriptor)
    assert o["ok"] and r["status"] == 0
    assert (r["kl"], r["kr"], r["m"], r["best"]) == (nkl, nkr, nm, o["best_count"])
    assert np.array_equal(bits(r["H"]), bits(o["H"])) and np.array_equal(canvas, want)
    # back to the reference's matcher
    canvas0, r0 = engine.stitchTwoImages(left, right)
    o0 = oracle.stitch_pair(left, right, seed=12345)
    assert r0["m"] == o0["stats"]["m"] > r["m"] and np.array_equal(canvas0, o0["canvas"])


def test_fold_and_batch_with_the_ratio_test_matcher(engine, oracle):
    pkg = load_pkg()
    views = load_synth().make_strip(n=3, w=640, h=400, seed=21)
    engine.set_match_mode(1, ratio=0.8)
    try:
        pano, log = engine.stitchAllImages(views)
        res = engine.stitchBatch([views[0], views[1]], [views[1], views[2]])
    finally:
        engine.set_match_mode(0)
    want = np.ascontiguousarray(views[0])
    for step, im in enumerate(views[1:]):
        o, nkl, nkr, nm, canvas = oracle_pair_knn(oracle, want, im, 0.8, 0)
        assert o["ok"] and log[step]["status"] == 0 and (log[step]["m"], log[step]["best"]) == (nm, o["best_count"])
        assert np.array_equal(bits(log[step]["H"]), bits(o["H"]))
        want = canvas
    assert np.array_equal(np.asarray(pano), want)
    results = res[0] if isinstance(res, tuple) else res
    for i, (l, r) in enumerate(((views[0], views[1]), (views[1], views[2]))):
        o, _, _, nm, _ = oracle_pair_knn(oracle, l, r, 0.8, 0)
        assert results[i]["status"] == 0 and results[i]["m"] == nm and np.array_equal(bits(results[i]["H"]), bits(o["H"]))
    with pytest.raises(pkg.PanoError):
        engine.set_match_mode(1, ratio=0.0)
