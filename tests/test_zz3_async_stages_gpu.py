"""GPU tier: the asynchronous forms of the stage calls (pano_detect_async / pano_match_async / pano_ransac_async;
SURVEY 8 b3) return what the blocking calls return, one pending call per context.  They reuse the worker mechanism of
pano_stitch_pair_async (tests/test_gpu_parity.py::test_async_pair_equals_blocking_pair) around the unchanged blocking
calls; written after the round's GPU budget was spent (first run on a B200 is the driver's), hence in a file that
sorts last."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_pkg

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


def test_async_stage_calls_equal_blocking_calls(engine, oracle, small_pair):
    pkg = load_pkg()
    left, right, _ = small_pair
    kl = engine.gpuHarrisCornerDetectorDetect(left)
    h = engine.gpuHarrisCornerDetectorDetectAsync(right)
    # one pending call per context
    opts = pkg.HarrisCornerOptions()
    n = C.c_int(0)
    L = pkg._Img(left)
    busy = engine.lib.pano_detect_async(engine.ctx, L.ptr, L.w, L.h, C.c_size_t(L.stride), 0, C.byref(opts), None, 0, C.byref(n), None)
    assert busy == pkg.PANO_ERR_BUSY
    kr = h.result()
    assert h.done() and np.array_equal(kr, engine.gpuHarrisCornerDetectorDetect(right)) and np.array_equal(kr, oracle.detect(right))
    hm = engine.gpuHarrisMatchKeyPointsAsync(kr, kl, right, left)
    m = hm.result()
    mb = engine.gpuHarrisMatchKeyPoints(kr, kl, right, left)
    assert len(m) == len(mb) > 0 and m.tobytes() == mb.tobytes()
    hr = engine.computeHomographyAsync(kr, kl, m)
    H, best, it = hr.result()
    Hb = engine.computeHomography(kr, kl, m)
    assert H is not None and np.array_equal(bits(H), bits(Hb)) and best > 0 and it >= 0
    o = oracle.ransac(kr, kl, m, seed=12345)
    assert np.array_equal(bits(H), bits(o["H"]))
    # a status of the underlying call arrives at completion: too few matches
    H2, _, _ = engine.computeHomographyAsync(kr, kl, m[:3]).result()
    assert H2 is None
    # the context is free again
    assert np.array_equal(engine.gpuHarrisCornerDetectorDetect(left), kl)
