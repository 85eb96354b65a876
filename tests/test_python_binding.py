"""CPU tier: the package's ctypes mirror of the reference's stage interface (`Engine.gpuHarrisCornerDetectorDetect`,
`gpuHarrisMatchKeyPoints`, `computeHomography`, the opt-in `matchKnn`, the asynchronous forms) marshals its arguments
and results correctly.  The C ABI underneath is the test-only CPU stand-in (tests/hostsim/abi_standin.cpp, on the oracle)
- NOT the product library, which refuses to create a context without a GPU (tests/test_abi.py) - so what is checked here
is the Python side: struct layouts, pointer / stride / count arguments, status handling, result slicing."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_pkg, load_synth


@pytest.fixture(scope="module")
def engine_on_standin():
    d = os.path.join(ROOT, "tests", "hostsim")
    so = os.path.join(d, "libabi_standin.so")
    oracle_so = os.path.join(ROOT, "oracle", "libpano_oracle.so")
    if not os.path.exists(oracle_so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libpano_oracle.so"], stdout=subprocess.DEVNULL)
    srcs = [os.path.join(d, "abi_standin.cpp"), os.path.join(ROOT, "include", "pano_b200.h"), oracle_so]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-I" + os.path.join(ROOT, "include"),
                               "-o", so, srcs[0], "-L" + os.path.join(ROOT, "oracle"), "-l:libpano_oracle.so",
                               "-Wl,-rpath,$ORIGIN/../../oracle"])
    pkg = load_pkg()
    lib = C.CDLL(so)
    lib.pano_last_error.restype = C.c_char_p
    e = pkg.Engine.__new__(pkg.Engine)          # the binding's methods on the stand-in library
    e.lib, e.ctx, e.device = lib, C.c_void_p(), 0
    assert lib.pano_create(0, C.c_uint32(12345), C.byref(e.ctx)) == 0
    yield e
    e.close()


@pytest.fixture(scope="module")
def pair():
    left, right, _ = load_synth().make_pair(320, 200, seed=9)
    return left, right


def test_knn_options_layout():
    pkg = load_pkg()
    assert C.sizeof(pkg.KnnOptions) == 16 and pkg.KnnOptions.ratio_.offset == 8
    o = pkg.KnnOptions()
    assert (o.patchSize_, o.descriptor_, o.ratio_) == (5, 0, 0.75)


def test_stage_calls_marshal_like_the_oracle(engine_on_standin, oracle, pair):
    e = engine_on_standin
    left, right = pair
    kl, kr = e.gpuHarrisCornerDetectorDetect(left), e.gpuHarrisCornerDetectorDetect(right)
    assert np.array_equal(kl, oracle.detect(left)) and np.array_equal(kr, oracle.detect(right))
    view = np.zeros((200, 330, 3), np.uint8)[:, :320]      # a non-contiguous row pitch
    view[:] = left
    assert np.array_equal(e.gpuHarrisCornerDetectorDetect(view), kl)
    m = e.gpuHarrisMatchKeyPoints(kr, kl, right, left)
    mo = oracle.match(kr, kl, right, left)
    assert len(m) > 50 and m.tobytes() == np.ascontiguousarray(mo).tobytes()
    m3 = e.gpuHarrisMatchKeyPoints(kr, kl, right, left, patchSize=3, maxSSDThresh=900.0, offset=4)
    assert m3.tobytes() == np.ascontiguousarray(oracle.match(kr, kl, right, left, patch=3, max_ssd=900.0, offset=4)).tobytes()
    H = e.computeHomography(kr, kl, m)
    o = oracle.ransac(kr, kl, mo, seed=12345)
    assert H is not None and np.array_equal(H.view(np.uint64), o["H"].view(np.uint64))
    assert e.computeHomography(kr, kl, m[:3]) is None       # too few matches: the reference's empty Mat


@pytest.mark.parametrize("descriptor,ratio,patch", [(0, 0.75, 5), (0, 0.9, 3), (1, 0.9, 5)])
def test_match_knn_marshals_like_the_checker(engine_on_standin, oracle, pair, descriptor, ratio, patch):
    e = engine_on_standin
    left, right = pair
    kl, kr = oracle.detect(left), oracle.detect(right)
    m, s = e.matchKnn(kr, kl, right, left, patchSize=patch, descriptor=descriptor, ratio=ratio)
    mo, so = oracle.match_knn(kr, kl, right, left, patch=patch, descriptor=descriptor, ratio=ratio)
    assert len(mo) > 3 and m.tobytes() == np.ascontiguousarray(mo).tobytes() and np.array_equal(s, so)


def test_match_knn_argument_errors_raise(engine_on_standin, pair):
    pkg = load_pkg()
    e = engine_on_standin
    left, right = pair
    k = np.array([[10, 10], [20, 20], [30, 30]], np.int32)
    for kw, code in ((dict(ratio=0.0), pkg.PANO_ERR_INVALID), (dict(ratio=1.5), pkg.PANO_ERR_INVALID),
                     (dict(descriptor=7), pkg.PANO_ERR_UNSUPPORTED), (dict(descriptor=1, patchSize=3), pkg.PANO_ERR_UNSUPPORTED)):
        with pytest.raises(pkg.PanoError) as err:
            e.matchKnn(k, k, right, left, **kw)
        assert err.value.status == code


def test_async_forms_marshal_like_the_blocking_calls(engine_on_standin, oracle, pair):
    pkg = load_pkg()
    e = engine_on_standin
    left, right = pair
    kl = e.gpuHarrisCornerDetectorDetect(left)
    h = e.gpuHarrisCornerDetectorDetectAsync(right)
    opts, n, L = pkg.HarrisCornerOptions(), C.c_int(0), pkg._Img(left)
    assert e.lib.pano_detect_async(e.ctx, L.ptr, L.w, L.h, C.c_size_t(L.stride), 0, C.byref(opts), None, 0, C.byref(n), None) == pkg.PANO_ERR_BUSY
    kr = h.result()
    assert h.done() and np.array_equal(kr, oracle.detect(right))
    m = e.gpuHarrisMatchKeyPointsAsync(kr, kl, right, left).result()
    assert m.tobytes() == e.gpuHarrisMatchKeyPoints(kr, kl, right, left).tobytes()
    H, best, it = e.computeHomographyAsync(kr, kl, m).result()
    assert np.array_equal(H.view(np.uint64), e.computeHomography(kr, kl, m).view(np.uint64)) and best > 0 and it >= 0
    H2, _, _ = e.computeHomographyAsync(kr, kl, m[:3]).result()
    assert H2 is None
