"""CPU tier: the WHOLE engine without a GPU.  tests/hostsim/build_emu_lib.py compiles the engine's own sources - the C
ABI, the host orchestration (stage sequencing, scratch buffers, read-backs, the async worker, replay planning and
re-runs) and every kernel, the tensor-core matcher included - with g++ against a fake, synchronous CUDA runtime
(tests/hostsim/fake_cuda) on the CPU emulation of the CUDA execution model (cuda_emu.hpp, tcgen05_emu.hpp); the only
source change is the mechanical rewrite of `kernel<<<...>>>(...)` launches.  The package's Python binding then drives
that library exactly as it drives libpano_b200.so on a B200, and the results must be the oracle's, bit for bit.
TEST INFRASTRUCTURE ONLY: the emulated library is never shipped and never loaded by the package (which refuses to work
without a GPU, tests/test_abi.py).  Sizes are small: a thread of the emulation is a fiber on one host core.
Also runs the GPU-tier test code of the pieces that were written after the round's GPU budget was spent (the opt-in
2-NN matcher, the asynchronous stage calls) against the emulated engine."""
import ctypes as C
import importlib.util
import os

import numpy as np
import pytest

from conftest import ROOT, load_pkg, load_synth


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


@pytest.fixture(scope="module")
def engine():
    os.environ["PANO_BATCH_LANES"] = "1"        # (the emulation runs one launch at a time)
    spec = importlib.util.spec_from_file_location("build_emu_lib", os.path.join(ROOT, "tests", "hostsim", "build_emu_lib.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    pkg = load_pkg()
    lib = C.CDLL(mod.build())
    lib.pano_last_error.restype = C.c_char_p
    lib.pano_version.restype = C.c_char_p
    lib.pano_kernel_launches.restype = C.c_uint64
    e = pkg.Engine.__new__(pkg.Engine)           # the binding's methods on the emulated library
    e.lib, e.ctx, e.device = lib, C.c_void_p(), 0
    assert lib.pano_create(0, C.c_uint32(12345), C.byref(e.ctx)) == 0
    yield e
    e.close()


@pytest.fixture(scope="module")
def pair():
    left, right, _ = load_synth().make_pair(320, 200, seed=9)
    return left, right


def test_emulated_engine_stitches_a_pair_like_the_reference(engine, oracle, pair):
    """pano_stitch_pair end to end with the reference's options (1000 RANSAC iterations): detection of both images, tensor-
    core matcher, shuffle replay on the side stream, DLT, scoring, canvas geometry, quad warp kernel, canvas fetch"""
    left, right = pair
    n0 = engine.kernel_launches()
    canvas, r = engine.stitchTwoImages(left, right)
    o = oracle.stitch_pair(left, right, seed=12345)
    assert r["status"] == 0 and o["status"] == 1
    assert (r["kl"], r["kr"], r["m"], r["best"]) == (o["stats"]["kl"], o["stats"]["kr"], o["stats"]["m"], o["stats"]["best"])
    assert np.array_equal(bits(r["H"]), bits(o["H"])) and np.array_equal(canvas, o["canvas"])
    assert engine.kernel_launches() - n0 > 20


def test_emulated_engine_stage_calls(engine, oracle, pair):
    pkg = load_pkg()
    left, right = pair
    kl, kr = engine.gpuHarrisCornerDetectorDetect(left), engine.gpuHarrisCornerDetectorDetect(right)
    assert np.array_equal(kl, oracle.detect(left)) and np.array_equal(kr, oracle.detect(right))
    assert np.array_equal(engine.gpuHarrisCornerDetectorDetect(left, nmsNeighborhood=5), oracle.detect(left, nbhd=5))
    assert np.array_equal(bits(engine.harrisResponse(right)), bits(oracle.harris_response(right)))
    mo = np.ascontiguousarray(oracle.match(kr, kl, right, left))
    for matcher in (0, 1):                                   # tensor-core matcher (on its host model), SIMT matcher
        engine.set_matcher(matcher)
        try:
            assert engine.gpuHarrisMatchKeyPoints(kr, kl, right, left).tobytes() == mo.tobytes()
            m3 = engine.gpuHarrisMatchKeyPoints(kr, kl, right, left, patchSize=3, maxSSDThresh=900.0, offset=2)
            assert m3.tobytes() == np.ascontiguousarray(oracle.match(kr, kl, right, left, patch=3, max_ssd=900.0, offset=2)).tobytes()
        finally:
            engine.set_matcher(0)
    ro = pkg.RansacOptions(numIterations_=80)
    d = engine.computeHomography(kr, kl, mo, options=ro, details=True)
    o = oracle.ransac(kr, kl, mo, iters=80, seed=12345)
    assert np.array_equal(d["samples"], o["samples"]) and np.array_equal(d["counts"], o["counts"])
    assert np.array_equal(bits(d["H"]), bits(o["H"])) and np.array_equal(d["inlier_mask"].astype(bool), o["inlier_mask"])
    bad = mo.copy()
    bad["trainIdx"][3] = len(kl) + 9                          # checked on the device, reported through the error word
    with pytest.raises(pkg.PanoError) as e:
        engine.computeHomography(kr, kl, bad, options=ro)
    assert e.value.status == pkg.PANO_ERR_INVALID
    assert np.array_equal(engine.warpOverlay(left, right, o["H"]), oracle.compose(left, right, o["H"]))
    M = np.array([[0.9, 0.05, 12.0], [-0.04, 1.1, -7.0], [1e-4, -2e-4, 1.0]])
    assert np.array_equal(engine.warpPerspective(right, M, (260, 180)), oracle.warp_perspective(right, M, (260, 180)))


def test_emulated_engine_fold_of_three_images(engine, oracle):
    """pano_stitch_fold (ref: stitchAllImages, src/serial/main.cpp:395-414): the second step detects on the panorama the
    first one produced.  150 RANSAC iterations per step (the emulation spends 12 ms per DLT hypothesis); the oracle's fold
    is assembled from its stage functions with the same count"""
    pkg = load_pkg()
    views = load_synth().make_strip(n=3, w=240, h=160, seed=21)
    iters = 150
    pano, log = engine.stitchAllImages(views, ransacOpts=pkg.RansacOptions(numIterations_=iters))
    want = np.ascontiguousarray(views[0])
    for step, im in enumerate(views[1:]):
        kl, kr = oracle.detect(want), oracle.detect(im)
        m = oracle.match(kr, kl, im, want)
        o = oracle.ransac(kr, kl, m, iters=iters, seed=12345)
        assert o["ok"] and log[step]["status"] == 0
        assert (log[step]["kl"], log[step]["kr"], log[step]["m"], log[step]["best"]) == (len(kl), len(kr), len(m), o["best_count"])
        assert np.array_equal(bits(log[step]["H"]), bits(o["H"]))
        want = oracle.compose(want, im, o["H"])
    assert np.array_equal(np.asarray(pano), want)


# ---- the GPU-tier test code of the late additions, against the emulated engine ------------------------------------------
def _load_gpu_test_module(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tests", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def knn_tests():
    return _load_gpu_test_module("test_zz4_knn_gpu")


@pytest.fixture(scope="module")
def scenes(oracle):
    """the GPU tier's 'small' scene (640 x 360) - the 1080p one is left to the B200"""
    left, right, _ = load_synth().make_pair(640, 360, seed=11)
    kl, kr = oracle.detect(left), oracle.detect(right)
    kl = np.concatenate([kl, [[0, 0], [1, 359], [639, 5]], kl[:9]]).astype(np.int32)
    kr = np.concatenate([[[2, 1]], kr, kr[3:6]]).astype(np.int32)
    return {"small": (left, right, kl, kr)}


def test_emulated_engine_knn_argument_checks(engine, scenes, knn_tests):
    knn_tests.test_knn_argument_checks(engine, scenes)


@pytest.mark.parametrize("matcher", ["simt", "tensor-core"])
def test_emulated_engine_knn_patch_ssd(engine, oracle, scenes, knn_tests, matcher):
    knn_tests.ssd_equals_checker(engine, oracle, scenes, knn_tests.SIMT if matcher == "simt" else knn_tests.TC, "small")


def test_emulated_engine_knn_binary(engine, oracle, scenes, knn_tests):
    left, right, kl, kr = scenes["small"]
    for ratio in (0.8, 1.0):
        knn_tests.check(engine, oracle, kr, kl, right, left, 1, ratio, min_matches=5)


@pytest.mark.parametrize("descriptor,matcher", [(1, "simt"), (0, "simt"), (0, "tensor-core")])
def test_emulated_engine_knn_edge_cases(engine, oracle, scenes, knn_tests, descriptor, matcher):
    knn_tests.edge_cases_equal_checker(engine, oracle, scenes, descriptor, knn_tests.SIMT if matcher == "simt" else knn_tests.TC)


def test_emulated_engine_knn_other_patch_sizes(engine, oracle, scenes, knn_tests):
    for patch in (1, 3):
        for matcher in (knn_tests.SIMT, knn_tests.TC):
            knn_tests.other_patch_size(engine, oracle, scenes, patch, matcher)


def test_emulated_engine_knn_at_ratio_one_is_the_reference_matcher(engine, scenes, knn_tests):
    small = {"mid": scenes["small"]}                          # (the GPU test's 1080p scene, replaced by the small one)
    knn_tests.test_knn_tensor_core_and_simt_agree_with_the_reference_matcher_at_ratio_one(engine, small)


def test_emulated_engine_async_stage_calls(engine, oracle, pair):
    """tests/test_zz3_async_stages_gpu.py's checks on a smaller pair with fewer iterations: the worker thread of the
    asynchronous forms drives the same emulated device"""
    pkg = load_pkg()
    left, right = pair
    kl = engine.gpuHarrisCornerDetectorDetect(left)
    h = engine.gpuHarrisCornerDetectorDetectAsync(right)
    opts, n, L = pkg.HarrisCornerOptions(), C.c_int(0), pkg._Img(left)
    busy = engine.lib.pano_detect_async(engine.ctx, L.ptr, L.w, L.h, C.c_size_t(L.stride), 0, C.byref(opts), None, 0, C.byref(n), None)
    assert busy == pkg.PANO_ERR_BUSY
    kr = h.result()
    assert h.done() and np.array_equal(kr, oracle.detect(right))
    m = engine.gpuHarrisMatchKeyPointsAsync(kr, kl, right, left).result()
    assert m.tobytes() == np.ascontiguousarray(oracle.match(kr, kl, right, left)).tobytes()
    ro = pkg.RansacOptions(numIterations_=60)
    H, best, it = engine.computeHomographyAsync(kr, kl, m, options=ro).result()
    o = oracle.ransac(kr, kl, m, iters=60, seed=12345)
    assert np.array_equal(bits(H), bits(o["H"])) and (best, it) == (o["best_count"], o["best_iter"])
    H2, _, _ = engine.computeHomographyAsync(kr, kl, m[:3]).result()
    assert H2 is None
    pair_async = engine.stitchTwoImagesAsync(left, right, ransacOpts=ro)
    r = pair_async.result()
    assert r["status"] == 0 and np.array_equal(bits(r["H"]), bits(o["H"]))
    assert np.array_equal(engine.getCanvas(), oracle.compose(left, right, o["H"]))


def oracle_pair_knn(oracle, left, right, ratio, descriptor, iters, seed=12345):
    """what the fused calls do in match mode 1, assembled from the checker's stage functions"""
    kl, kr = oracle.detect(left), oracle.detect(right)
    m, _ = oracle.match_knn(kr, kl, right, left, descriptor=descriptor, ratio=ratio)
    if len(m) < 4:
        return None, len(kl), len(kr), len(m), None
    o = oracle.ransac(kr, kl, m, iters=iters, seed=seed)
    return o, len(kl), len(kr), len(m), (oracle.compose(left, right, o["H"]) if o["ok"] else None)


@pytest.mark.parametrize("descriptor,ratio", [(0, 0.75), (1, 0.9)])
def test_emulated_engine_pair_with_the_ratio_test_matcher(engine, oracle, pair, descriptor, ratio):
    """pano_set_match_mode(ctx, 1, ...): the fused pair runs detection, the 2-NN matcher (tensor-core top-2 epilogue on its
    model, or the binary path), RANSAC on the ratio-tested matches, and the warp - and goes back to the reference's
    matcher afterwards"""
    pkg = load_pkg()
    left, right = pair
    iters = 120
    ro = pkg.RansacOptions(numIterations_=iters)
    engine.set_match_mode(1, ratio=ratio, descriptor=descriptor)
    try:
        canvas, r = engine.stitchTwoImages(left, right, ransacOpts=ro)
    finally:
        engine.set_match_mode(0)
    o, nkl, nkr, nm, want = oracle_pair_knn(oracle, left, right, ratio, descriptor, iters)
    assert o is not None and o["ok"] and r["status"] == 0
    assert (r["kl"], r["kr"], r["m"], r["best"]) == (nkl, nkr, nm, o["best_count"])
    assert np.array_equal(bits(r["H"]), bits(o["H"])) and np.array_equal(canvas, want)
    ref_m = len(oracle.match(oracle.detect(right), oracle.detect(left), right, left))
    assert r["m"] < ref_m                                             # the ratio test removed matches
    _, r0 = engine.stitchTwoImages(left, right, ransacOpts=ro, fetch=False)
    assert r0["m"] == ref_m                                           # mode 0 again: the reference's matcher
    with pytest.raises(pkg.PanoError):
        engine.set_match_mode(1, ratio=1.5)
    engine.set_match_mode(1, descriptor=pkg.KNN_BINARY)
    try:
        with pytest.raises(pkg.PanoError) as e:                        # the binary descriptor is defined on 5 x 5 patches
            engine.stitchTwoImages(left, right, harrisOpts=pkg.HarrisCornerOptions(patchSize_=3))
        assert e.value.status == pkg.PANO_ERR_UNSUPPORTED
    finally:
        engine.set_match_mode(0)


def test_emulated_engine_batch_of_pairs_in_both_match_modes(engine, oracle):
    """pano_stitch_batch (two lane threads taking turns on the emulated device; slots, the A / B pipeline, the shared
    mt19937 stream, canvases written to host buffers) = the pairs one by one, with the reference's matcher and with the
    opt-in ratio-test matcher handed down to the slots"""
    pkg = load_pkg()
    os.environ["PANO_BATCH_LANES"] = "2"
    views = load_synth().make_strip(n=3, w=240, h=160, seed=21)
    iters = 100
    ro = pkg.RansacOptions(numIterations_=iters)
    lefts, rights = [views[0], views[1]], [views[1], views[2]]
    cap = 3 * 700 * 300
    outs = [np.zeros(cap, np.uint8) for _ in lefts]
    res, _ = engine.stitchBatch(lefts, rights, ransacOpts=ro, canvases_out=outs)
    for i, (l, r) in enumerate(zip(lefts, rights)):
        kl, kr = oracle.detect(l), oracle.detect(r)
        o = oracle.ransac(kr, kl, oracle.match(kr, kl, r, l), iters=iters, seed=12345)
        want = oracle.compose(l, r, o["H"])
        assert res[i]["status"] == 0 and np.array_equal(bits(res[i]["H"]), bits(o["H"]))
        assert np.array_equal(outs[i][:want.size].reshape(want.shape), want)
    engine.set_match_mode(1, ratio=0.8)
    try:
        res, _ = engine.stitchBatch(lefts, rights, ransacOpts=ro)
    finally:
        engine.set_match_mode(0)
        os.environ["PANO_BATCH_LANES"] = "1"
    for i, (l, r) in enumerate(zip(lefts, rights)):
        o, _, _, nm, _ = oracle_pair_knn(oracle, l, r, 0.8, 0, iters)
        assert res[i]["status"] == 0 and res[i]["m"] == nm and np.array_equal(bits(res[i]["H"]), bits(o["H"]))


def test_emulated_gpu_stitching_executable_chain_mode(engine, oracle, tmp_path):
    """host/gpu_stitching.cpp itself (with reader and image codecs) linked against the emulated library and the fake CUDA
    runtime: PANO_MODE=chain over two "devices" (worker threads, per-device contexts, bands into the host canvas; the
    check of tests/test_zz2_chain_cli_gpu.py) gives the oracle's chain panorama (the default fold of the executable is covered by tests/test_cli.py on the GPU)"""
    import subprocess
    cv2 = pytest.importorskip("cv2")
    exe = os.path.join(ROOT, "tests", "hostsim", "gpu_stitching_emu")
    assert os.path.exists(exe)                                   # built together with the emulated library (fixture `engine`)
    views = load_synth().make_strip(n=3, w=240, h=160, seed=21)
    paths = []
    for i, v in enumerate(views):
        p = str(tmp_path / ("v%d.ppm" % i))
        assert cv2.imwrite(p, v)
        paths.append(p)
    pano, pair_H = oracle.stitch_chain(views, seed=12345)
    assert all(H is not None for H in pair_H)
    out = str(tmp_path / "chain.png")
    r = subprocess.run([exe] + paths + ["--out", out], capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, PANO_MODE="chain", PANO_GPUS="2", PANO_EMU_DEVICES="4"))
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Chain mode: 3 images, adjacent pairs sharded over 2 GPU(s)" in r.stdout and "(GPU 1)" in r.stdout
    for needle in ("Harris Corner Detection (GPU): ", "RANSAC Homography Estimation (GPU): ", "Image Stitching: ",
                   "Total Stitching Process: ", "Stitched result saved to " + out, "Total Execution Time: "):
        assert needle in r.stdout, needle
    assert np.array_equal(cv2.imread(out), pano)


def test_reference_gpu_main_on_the_emulated_engine(engine, tmp_path):
    """SURVEY 8 b2 without a GPU: the reference's own GPU executable - src/gpu/main.cpp compiled UNMODIFIED from
    /root/reference - with its four .cu stage files replaced by the maintainer-side binding
    (examples/reference_shim/pano_b200_shim.cpp -> C ABI), linked against the EMULATED engine library: detection, matching
    (tensor-core matcher on its model) and RANSAC are the engine's own code, geometry / warp / overlay the reference's.
    Its panorama must be the one the reference's serial code produces for the same seed.  (tests/test_reference_shim.py
    does this on a stand-in of the ABI, tests/test_zz1_reference_gpu_main.py on a B200.)"""
    import subprocess
    cv2 = pytest.importorskip("cv2")
    ref_root = "/root/reference"
    if not os.path.exists(os.path.join(ref_root, "src", "gpu", "main.cpp")):
        pytest.skip("/root/reference not present: nothing to compile")
    from oracle import ref as refmod
    if not refmod.available():
        pytest.skip("oracle/_ref not built")
    o = os.path.join(ROOT, "oracle")
    hs = os.path.join(ROOT, "tests", "hostsim")
    exe = str(tmp_path / "gpu_stitching_refmain_emu")
    subprocess.check_call(
        ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-w", "-I" + os.path.join(o, "cvshim"), "-I" + ref_root + "/src",
         "-I" + ref_root + "/src/reader", "-I" + ref_root + "/src/gpu", "-I" + os.path.join(ROOT, "include"), "-o", exe,
         ref_root + "/src/gpu/main.cpp", ref_root + "/src/reader/reader.cpp",
         os.path.join(ROOT, "examples", "reference_shim", "pano_b200_shim.cpp"), os.path.join(o, "cvshim", "cvshim.cpp"),
         os.path.join(hs, "libpano_b200_emu.so"), "-Wl,-rpath," + hs, "-lpthread"])
    left, right, _ = load_synth().make_pair(320, 200, seed=9)
    a, b, out = str(tmp_path / "a.ppm"), str(tmp_path / "b.ppm"), str(tmp_path / "pano.ppm")
    assert cv2.imwrite(a, left) and cv2.imwrite(b, right)
    r = subprocess.run([exe, a, b, "--out", out], capture_output=True, text=True, timeout=900, env=dict(os.environ, PANO_SEED="7"))
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Harris Corner Matching (GPU):" in r.stdout and "RANSAC Homography Estimation (GPU):" in r.stdout
    assert "falling back" not in r.stderr
    ref = refmod.Reference().stitch_pair(left, right, seed=7)
    assert ref["status"] == 1 and np.array_equal(cv2.imread(out), ref["canvas"])


def test_emulated_engine_survives_a_matcher_pipeline_timeout(engine, tmp_path):
    """The tensor-core matcher bounds every pipeline wait; a wait that gives up raises the context's error word.  The host
    must then fail THAT call loudly (never emit matches from an aborted kernel), clear the word, stop using the
    tensor-core matcher in this process and answer the next call with the SIMT matcher.  The device has never taken this
    path in a test; here the TMA model drops one tile load, its stage barrier never completes and the (bounded) waits
    give up (tests/hostsim/tcgen05_emu.hpp fault injection).  Runs in a child process because the switch to the SIMT
    matcher is process-wide."""
    import subprocess
    import sys
    script = tmp_path / "abort_child.py"
    script.write_text("""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from conftest import load_pkg, load_synth
from oracle.oracle import Oracle
pkg, o = load_pkg(), Oracle()
lib = C.CDLL(%r)
lib.pano_last_error.restype = C.c_char_p
e = pkg.Engine.__new__(pkg.Engine)
e.lib, e.ctx, e.device = lib, C.c_void_p(), 0
assert lib.pano_create(0, C.c_uint32(12345), C.byref(e.ctx)) == 0
left, right, _ = load_synth().make_pair(320, 200, seed=9)
kl, kr = o.detect(left), o.detect(right)
want = np.ascontiguousarray(o.match(kr, kl, right, left)).tobytes()
assert e.gpuHarrisMatchKeyPoints(kr, kl, right, left).tobytes() == want          # tensor-core matcher (on its model)
os.environ["PANO_EMU_TC_DROP_LOAD"] = "1"
try:
    e.gpuHarrisMatchKeyPoints(kr, kl, right, left)
    raise SystemExit("the aborted call returned matches")
except pkg.PanoError as err:
    assert err.status == pkg.PANO_ERR_CUDA and "tensor-core matcher pipeline timed out" in str(err), str(err)
del os.environ["PANO_EMU_TC_DROP_LOAD"]
n0 = e.kernel_launches()
assert e.gpuHarrisMatchKeyPoints(kr, kl, right, left).tobytes() == want          # the SIMT matcher has taken over
canvas, r = e.stitchTwoImages(left, right, ransacOpts=pkg.RansacOptions(numIterations_=40))
assert r["status"] == 0 and r["m"] == len(o.match(kr, kl, right, left))
os.environ["PANO_EMU_TC_DROP_LOAD"] = "0"                                        # no tensor-core kernel runs any more
assert e.gpuHarrisMatchKeyPoints(kr, kl, right, left).tobytes() == want
print("OK")
""" % (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "hostsim", "libpano_b200_emu.so")))
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), (r.stdout[-500:], r.stderr[-1500:])


def test_emulated_engine_survives_every_failed_allocation(engine, oracle):
    """memory exhaustion at every allocation of a pair in turn (fault injection in the fake runtime: the n-th cudaMalloc /
    cudaMallocHost fails): the call must return PANO_ERR_CUDA - never crash, never return a wrong result - and the same
    context must stitch the pair correctly right afterwards (grow-only buffers left in a consistent state)"""
    pkg = load_pkg()
    lib = engine.lib
    left, right, _ = load_synth().make_pair(200, 140, seed=9)
    iters = 12
    ro = pkg.RansacOptions(numIterations_=iters)
    kl, kr = oracle.detect(left), oracle.detect(right)
    o = oracle.ransac(kr, kl, oracle.match(kr, kl, right, left), iters=iters, seed=12345)
    want = oracle.compose(left, right, o["H"])
    n = handled = 0
    try:
        while True:
            e = pkg.Engine.__new__(pkg.Engine)
            e.lib, e.ctx, e.device = lib, C.c_void_p(), 0
            lib.pano_emu_fail_malloc(C.c_long(n))
            if lib.pano_create(0, C.c_uint32(12345), C.byref(e.ctx)) != 0:       # the failure hit the context's own buffers
                e.ctx = None
                lib.pano_emu_fail_malloc(C.c_long(-1))
                handled += 1
                n += 1
                continue
            failed = False
            try:
                e.stitchTwoImages(left, right, ransacOpts=ro)
            except pkg.PanoError as err:
                failed = True
                assert err.status == pkg.PANO_ERR_CUDA
            lib.pano_emu_fail_malloc(C.c_long(-1))
            canvas, r = e.stitchTwoImages(left, right, ransacOpts=ro)
            assert r["status"] == 0 and np.array_equal(bits(r["H"]), bits(o["H"])) and np.array_equal(canvas, want), n
            e.close()
            if not failed:
                break
            handled += 1
            n += 1
            assert n < 200
    finally:
        lib.pano_emu_fail_malloc(C.c_long(-1))
    assert handled >= 20          # a pair allocates a few dozen buffers on a fresh context: every one of them was failed once


def test_emulated_engine_batch_survives_failed_allocations_in_lane_threads(engine, oracle):
    """the same for pano_stitch_batch with two lane threads: an allocation that fails inside a lane (slot contexts, upload
    and scratch buffers, replay plans) must come back as PANO_ERR_CUDA of the batch call - an exception leaving a lane's
    std::thread would terminate the host process - and the next batch on the same context must be right.  Every seventh
    allocation index up to the first success (the full sweep - 116 indices for three pairs - was run once: all handled)."""
    pkg = load_pkg()
    lib = engine.lib
    views = load_synth().make_strip(n=3, w=200, h=140, seed=21)
    lefts, rights = list(views[:2]), list(views[1:])
    iters = 10
    ro = pkg.RansacOptions(numIterations_=iters)
    want = []
    for l, r in zip(lefts, rights):
        kl, kr = oracle.detect(l), oracle.detect(r)
        want.append(oracle.ransac(kr, kl, oracle.match(kr, kl, r, l), iters=iters, seed=12345)["H"])
    os.environ["PANO_BATCH_LANES"] = "2"
    n = handled = 0
    try:
        while n < 400:
            e = pkg.Engine.__new__(pkg.Engine)
            e.lib, e.ctx, e.device = lib, C.c_void_p(), 0
            assert lib.pano_create(0, C.c_uint32(12345), C.byref(e.ctx)) == 0
            lib.pano_emu_fail_malloc(C.c_long(n))
            failed = False
            try:
                e.stitchBatch(lefts, rights, ransacOpts=ro)
            except pkg.PanoError as err:
                failed = True
                assert err.status == pkg.PANO_ERR_CUDA
            lib.pano_emu_fail_malloc(C.c_long(-1))
            res, _ = e.stitchBatch(lefts, rights, ransacOpts=ro)
            for i in range(2):
                assert res[i]["status"] == 0 and np.array_equal(bits(res[i]["H"]), bits(want[i])), (n, i)
            e.close()
            if not failed:
                break
            handled += 1
            n += 7
    finally:
        lib.pano_emu_fail_malloc(C.c_long(-1))
        os.environ["PANO_BATCH_LANES"] = "1"
    assert handled >= 8


def test_emulated_engine_rejects_bad_arguments_without_touching_memory(engine):
    """every entry point of the C ABI with arguments that must be refused: a status, never a crash (under
    tools/emu_sanitize.sh also: never an out-of-bounds access), and the context works afterwards"""
    pkg = load_pkg()
    lib, ctx = engine.lib, engine.ctx
    INVALID, UNSUPPORTED, TOO_FEW = pkg.PANO_ERR_INVALID, pkg.PANO_ERR_UNSUPPORTED, pkg.PANO_ERR_TOO_FEW_MATCHES
    img = np.ascontiguousarray(load_synth().make_pair(96, 64, seed=3)[0])
    ip, w, h, st = img.ctypes.data_as(C.c_void_p), 96, 64, C.c_size_t(img.strides[0])
    ho, ro, ko = pkg.HarrisCornerOptions(), pkg.RansacOptions(), pkg.KnnOptions()
    n = C.c_int(0)
    xy = np.zeros((64, 2), np.int32)
    xp = xy.ctypes.data_as(C.c_void_p)
    m = np.zeros(8, pkg.MATCH_DTYPE)
    mp = m.ctypes.data_as(C.c_void_p)
    H = np.eye(3)
    Hp = H.ctypes.data_as(C.c_void_p)
    res = pkg.PairResult()
    info = pkg.CanvasInfo()
    canvas = np.zeros(1 << 16, np.uint8)
    cp = canvas.ctypes.data_as(C.c_void_p)

    def opts(**kw):
        o = pkg.HarrisCornerOptions()
        for k, v in kw.items():
            setattr(o, k, v)
        return o
    cases = [
        (lib.pano_detect(ctx, None, w, h, st, 0, C.byref(ho), xp, 64, C.byref(n)), INVALID),
        (lib.pano_detect(ctx, ip, 0, h, st, 0, C.byref(ho), xp, 64, C.byref(n)), INVALID),
        (lib.pano_detect(ctx, ip, w, h, C.c_size_t(3 * w - 1), 0, C.byref(ho), xp, 64, C.byref(n)), INVALID),
        (lib.pano_detect(ctx, ip, w, h, st, 0, None, xp, 64, C.byref(n)), INVALID),
        (lib.pano_detect(ctx, ip, w, h, st, 0, C.byref(ho), xp, 64, None), INVALID),
        (lib.pano_detect(ctx, ip, w, h, st, 0, C.byref(opts(nmsNeighborhood_=4)), xp, 64, C.byref(n)), UNSUPPORTED),
        (lib.pano_detect(ctx, ip, w, h, st, 0, C.byref(opts(patchSize_=7)), xp, 64, C.byref(n)), UNSUPPORTED),
        (lib.pano_detect(None, ip, w, h, st, 0, C.byref(ho), xp, 64, C.byref(n)), INVALID),
        (lib.pano_harris_response(ctx, ip, w, h, st, 0, C.c_double(0.04), None), INVALID),
        (lib.pano_convolve_f64(ctx, cp, 8, 8, cp, 4, 0, cp), INVALID),
        (lib.pano_match(ctx, xp, -1, xp, 4, ip, w, h, st, ip, w, h, st, 0, C.byref(ho), 0, mp, 8, C.byref(n)), INVALID),
        (lib.pano_match(ctx, None, 4, xp, 4, ip, w, h, st, ip, w, h, st, 0, C.byref(ho), 0, mp, 8, C.byref(n)), INVALID),
        (lib.pano_match(ctx, xp, 4, xp, 4, ip, w, h, st, None, w, h, st, 0, C.byref(ho), 0, mp, 8, C.byref(n)), INVALID),
        (lib.pano_match_knn(ctx, xp, 4, xp, 4, ip, w, h, st, ip, w, h, st, 0, None, mp, None, 8, C.byref(n)), INVALID),
        (lib.pano_match_knn(ctx, xp, 4, xp, 4, ip, w, h, st, ip, w, h, st, 0, C.byref(ko), mp, None, -1, C.byref(n)), INVALID),
        (lib.pano_ransac(ctx, xp, 8, xp, 8, mp, 8, 0, None, Hp, None, None, None, None, None), INVALID),
        (lib.pano_ransac(ctx, xp, 8, xp, 8, mp, 8, 0, C.byref(pkg.RansacOptions(numSamples_=3)), Hp, None, None, None, None, None), UNSUPPORTED),
        (lib.pano_ransac(ctx, xp, 8, xp, 8, mp, 3, 0, C.byref(ro), Hp, None, None, None, None, None), TOO_FEW),
        (lib.pano_ransac(ctx, xp, 8, xp, 8, mp, 8, 0, C.byref(pkg.RansacOptions(numIterations_=0)), Hp, None, None, None, None, None), TOO_FEW),
        (lib.pano_canvas_geometry(0, h, w, h, Hp, C.byref(info)), INVALID),
        (lib.pano_canvas_geometry(w, h, w, h, None, C.byref(info)), INVALID),
        (lib.pano_warp_overlay(ctx, ip, w, h, st, ip, w, h, st, 0, None, cp, C.c_size_t(3 * w), C.c_size_t(canvas.nbytes), C.byref(info)), INVALID),
        (lib.pano_stitch_pair(ctx, ip, w, h, st, None, w, h, st, 0, C.byref(ho), C.byref(ro), C.byref(res)), INVALID),
        (lib.pano_stitch_pair(ctx, ip, w, h, st, ip, w, h, st, 0, C.byref(ho), C.byref(ro), None), INVALID),
        (lib.pano_stitch_pair(ctx, ip, w, h, st, ip, w, h, st, 0, C.byref(ho), C.byref(pkg.RansacOptions(numSamples_=5)), C.byref(res)), UNSUPPORTED),
        (lib.pano_set_match_mode(ctx, 2, C.c_double(0.75), 0), INVALID),
        (lib.pano_set_match_mode(ctx, 1, C.c_double(-1.0), 0), INVALID),
        (lib.pano_set_match_mode(ctx, 1, C.c_double(0.75), 9), UNSUPPORTED),
        (lib.pano_pair_query(ctx), INVALID),                       # nothing pending
        (lib.pano_pair_wait(ctx), INVALID),
        (lib.pano_get_profile(ctx, None, None, 0) >= 0, True),
    ]
    for i, (got, want) in enumerate(cases):
        assert got == want, (i, got, want)
    # a flat image pair: the reference's failure cases as statuses, not errors
    flat = np.full((64, 96, 3), 50, np.uint8)
    canvas_, r = engine.stitchTwoImages(flat, flat)
    assert canvas_ is None and r["status"] == pkg.PANO_ERR_NO_MATCHES
    # and the context is intact
    k = engine.gpuHarrisCornerDetectorDetect(img)
    assert len(k) > 0



def test_whole_engine_fuzz_tool_short_run():
    """tools/fuzz_emu_engine.py (random scenes, options, seeds, engine settings and calls through the emulated library,
    against the composition of the oracle's stage functions): a short run inside the suite; the long runs are logged in
    profiles/r02_fuzz_emulated_engine_vs_oracle_build_container.jsonl"""
    import json
    import subprocess
    import sys
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_emu_engine.py"), "--cases", "24", "--seed", "77"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, (p.stdout[-1500:], p.stderr[-1500:])
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["ok"] and line["cases"] == 24 and line["pair"] + line["homography_only"] + line["fold"] + line["batch"] > 12
